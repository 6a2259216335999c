"""Oracle parity of the EXACT pipeline bench.py times, at cfg2's full size (32 pairs x 1000 hypotheses x 2000
correspondences): `drb_sample_sets` -> `drb_solve_e5` -> the scorer the pipelined service picks (tensor cores) ->
`drb_best_finalize`, against the fp64 CPU oracle fed the SAME minimal samples:

    idx (ours) -> oracle.nister.five_point(matches[idx].double())      nister.py:69-408
               -> oracle.scoring.msac_score(matches.double(), ...)      msac_score.py:12-55
               -> arg-max over the genuine models                       ransac.py:114

Per pair the outcome is classified:
    same    the same winning hypothesis (`best_id // 10`)
    tie     another hypothesis whose score is within 1e-4 relative of the oracle's best (on noise-free synthetic pairs
            every all-inlier sample scores the same to ~1e-6, and which of them is "best" is decided by rounding -- in
            the reference too)
    higher  our winner scores MORE than the oracle's best by over 1e-4 (and the fp64 oracle, scoring OUR model,
            confirms that score): the fp32 solver's model of an ill-conditioned sample is ~1e-5 away from the exact
            minimal solution and happens to fit the consensus better.  Seen on at most one pair of the 32, by 1.5e-4
    worse   our winner scores less than the oracle's best by over 1e-4: the solver lost the best model
The bar: no `worse`, at most two `higher` and none beyond 1e-3.  The counts and every pair's numbers are written to
gpurun_out/r2_pipeline_parity.json.

Second bar (scores on identical inputs): the score our scorer gives OUR winning model equals the fp64 oracle's
score of that same model within 1e-4 relative -- for the tensor-core scorer and for the FP32 work-queue kernel."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
DEV = "cuda"


@pytest.fixture(scope="module")
def cfg2():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    sys.path.insert(0, ROOT)
    import bench
    from helpers import trace_constraint_residual
    from oracle import nister, scoring
    from differentiable_ransac_b200 import engine

    B, K, N = 32, 1000, 2000
    matches, logits, thr, _ = bench.make_inputs(B, N, seed=1234)
    runs = {}
    scorers = [engine.service_scorer(torch.device(DEV), B), "stream"]
    for sc in scorers:
        o = engine.ransac_e5_test(matches.to(DEV), logits.to(DEV), K, thr.to(DEV), seed=42, offset=3,
                                  want_scores=True, scorer=sc)
        runs[sc] = {k: v.cpu() for k, v in o.items() if torch.is_tensor(v)}
    idx = runs[scorers[0]]["idx"].long()
    assert torch.equal(idx, runs["stream"]["idx"].long())          # same (seed, offset) -> same minimal samples
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    oracle = []
    for b in range(B):
        m64 = nister.five_point(matches[b][idx[b]].double())        # [K*10,3,3]
        s64, _ = scoring.msac_score(matches[b].double(), m64, float(thr[b]))
        genuine = trace_constraint_residual(m64) < 1e-8
        s64 = torch.where(genuine, s64, torch.full_like(s64, -1.0))
        oracle.append(dict(best=int(s64.argmax()), score=float(s64.max()), n_genuine=int(genuine.sum())))
    return dict(B=B, K=K, N=N, matches=matches, thr=thr, runs=runs, scorers=scorers, oracle=oracle)


def test_the_benched_pipeline_picks_the_oracles_winner_or_a_tie(cfg2):
    from oracle import scoring

    report = {}
    for sc in cfg2["scorers"]:
        o = cfg2["runs"][sc]
        same = ties = worse = higher = 0
        worst_tie, worst_model_score, worst_higher = 0.0, 0.0, 0.0
        detail = []
        for b in range(cfg2["B"]):
            ref = cfg2["oracle"][b]
            ours_hyp, ours_score = int(o["best_hyp"][b]), float(o["best_score"][b])
            rel = abs(ours_score - ref["score"]) / ref["score"]
            if ours_hyp == ref["best"] // 10:
                same += 1
            elif rel <= 1e-4:
                ties += 1
                worst_tie = max(worst_tie, rel)
            elif ours_score > ref["score"]:
                higher += 1
                worst_higher = max(worst_higher, rel)
            else:
                worse += 1
            # identical inputs: OUR winning model scored by the fp64 oracle
            s_model, _ = scoring.msac_score(cfg2["matches"][b].double(), o["best_model"][b].double()[None],
                                            float(cfg2["thr"][b]))
            rel_m = abs(float(s_model[0]) - ours_score) / max(1.0, float(s_model[0]))
            worst_model_score = max(worst_model_score, rel_m)
            detail.append(dict(pair=b, ours_hyp=ours_hyp, oracle_hyp=ref["best"] // 10, ours=ours_score,
                               oracle=ref["score"], rel=rel, rel_score_of_our_model=rel_m))
        report[sc] = dict(pairs=cfg2["B"], same_best_hypothesis=same, ties_within_1e_4=ties, higher_than_oracle=higher,
                          worse=worse, worst_tie_rel=worst_tie, worst_higher_rel=worst_higher,
                          worst_rel_score_on_identical_model=worst_model_score, detail=detail)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_pipeline_parity.json"), "w") as f:
        json.dump(dict(workload="cfg2: 32 x 1000 x 2000, sample_sets(seed=42, offset=3)", report=report), f, indent=1)
    for sc, r in report.items():
        print(sc, {k: v for k, v in r.items() if k != "detail"})
        assert r["worse"] == 0, (sc, [d for d in r["detail"] if d["ours_hyp"] != d["oracle_hyp"] and d["rel"] > 1e-4])
        assert r["higher_than_oracle"] <= 2 and r["worst_higher_rel"] <= 1e-3, (sc, r["higher_than_oracle"])
        assert r["worst_rel_score_on_identical_model"] <= 1e-4, (sc, r["worst_rel_score_on_identical_model"])


def test_both_scorers_agree_on_the_winner(cfg2):
    """The tensor-core scorer and the FP32 kernel see the same models: same winner, or winners whose scores agree
    to 1e-4 (the tie again)."""
    a, b = (cfg2["runs"][s] for s in cfg2["scorers"])
    for p in range(cfg2["B"]):
        if int(a["best_id"][p]) != int(b["best_id"][p]):
            assert abs(float(a["best_score"][p]) - float(b["best_score"][p])) <= 1e-4 * float(b["best_score"][p])
