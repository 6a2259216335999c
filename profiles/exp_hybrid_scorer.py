"""Feasibility: score part of the models on the tensor cores and the rest on the CUDA cores AT THE SAME TIME.

The tensor-core scorer leaves 44 % of the issue slots and 63 % of the FMA pipe idle and, built
slim, 16 K registers and ~50 KB of shared memory per SM: room for four one-warp CTAs of the FP32 block scorer.  Does a
launch of `drb_score_msac` over the LAST (1 - f) of every pair's models, on a second stream, hide under a launch of
`drb_score_msac_tc` over the first f?  cfg2 model list, CUDA events around both streams, L2 flushed.

    python profiles/exp_hybrid_scorer.py [tc variant]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from differentiable_ransac_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
tc = sys.argv[1] if len(sys.argv) > 1 else "tc_bf16p_s"
B, K, N = 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
_, _, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
M = cm.shape[1]
s_main, s_side = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1.0)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


best = torch.zeros(B, dtype=torch.int64, device=dev)


def whole(kernel):
    best.zero_()
    ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, best=best, kernel=kernel)


ref_best = None
for kern in (tc, "block", "stream"):
    med, mn = timed(lambda: whole(kern))
    print(json.dumps(dict(what="whole list", kernel=kern, ms_median=round(med, 5), ms_min=round(mn, 5))), flush=True)
whole(tc)
torch.cuda.synchronize()
ref_best = best.clone()

cur = torch.cuda.current_stream()
for frac in (0.9, 0.85, 0.8, 0.75, 0.7):
    # the first M1 models of every pair on the tensor cores, the rest on the CUDA cores (M1 a multiple of 128)
    mean_cnt = float(cc.float().mean())
    M1 = int(frac * mean_cnt) // 128 * 128
    cm_a, cid_a = cm[:, :M1].contiguous(), cid[:, :M1].contiguous()
    cm_b, cid_b = cm[:, M1:].contiguous(), cid[:, M1:].contiguous()
    cc_a, cc_b = cc.clamp(max=M1).contiguous(), (cc - M1).clamp(min=0).contiguous()
    for side in ("block", "stream"):
        def both():
            best.zero_()
            s_main.wait_stream(cur)
            s_side.wait_stream(cur)
            with torch.cuda.stream(s_side):
                ops.score_msac(m, cm_b, thr, count=cc_b, ids=cid_b, want_scores=False, best=best, kernel=side)
            with torch.cuda.stream(s_main):
                ops.score_msac(m, cm_a, thr, count=cc_a, ids=cid_a, want_scores=False, best=best, kernel=tc)
            cur.wait_stream(s_main)
            cur.wait_stream(s_side)
        med, mn = timed(both)
        both()
        torch.cuda.synchronize()
        same = int((best == ref_best).sum())
        print(json.dumps(dict(what="split", tc=tc, side=side, frac=frac, M1=M1, models_tc=int(cc_a.sum()),
                              models_side=int(cc_b.sum()), ms_median=round(med, 5), ms_min=round(mn, 5),
                              same_best_as_whole=same)), flush=True)
