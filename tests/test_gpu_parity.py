"""GPU parity tests: the CUDA path (through the C ABI, via ops/engine) against the CPU
oracle and the reference-generated golden fixtures.  Run on the B200 box: -m gpu."""
import math

import pytest
import torch

from helpers import match_up_to_sign, trace_constraint_residual, unit

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def drb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from differentiable_ransac_b200 import engine, ops, synth

    class NS:
        pass

    ns = NS()
    ns.ops, ns.engine, ns.synth = ops, engine, synth
    return ns


# ---- a1/a2 sampler -------------------------------------------------------------------------------
@pytest.mark.parametrize("regime", ["L0", "L1"])
def test_sampler_injected_noise_matches_reference(drb, golden, regime):
    g = golden(f"sampler_{regime}")
    idx, lse, sel_key, _ = drb.ops.sample(g["logits"][None].to(DEV), 12, 5, 1.0, noise=g["noise"][None].to(DEV),
                                          want_lse=True)
    assert torch.equal(idx[0].cpu().long(), g["idx"])                       # bit-exact top-s, ascending order
    keys = g["logits"][None, :] + g["noise"]
    assert torch.allclose(lse[0].cpu(), torch.logsumexp(keys, -1), rtol=1e-5, atol=1e-5)
    assert torch.equal(sel_key[0].cpu(), torch.gather(keys, 1, g["idx"]))
    assert torch.equal(g["matches"][idx[0].cpu().long()], g["minimal"])     # the gather of ransac.py:64-65


@pytest.mark.parametrize("s,N,K", [(3, 777, 33), (5, 2000, 64), (7, 1001, 17), (8, 4096, 40)])
def test_sampler_topk_vs_oracle(drb, s, N, K):
    from oracle import sampler as osamp
    B = 3
    logits = drb.synth.logits_regime(B, N, "L0", seed=s)
    noise = drb.synth.gumbel_noise((B, K, N), seed=10 + s)
    for tau in (1.0, 0.5):
        idx, lse, _, _ = drb.ops.sample(logits.to(DEV), K, s, tau, noise=noise.to(DEV), want_lse=True)
        for b in range(B):
            _, y_soft, ref_idx = osamp.sample(logits[b], noise[b], s, tau)
            assert torch.equal(idx[b].cpu().long(), ref_idx)
            keys = osamp.gumbel_keys(logits[b], noise[b], tau)
            assert torch.allclose(lse[b].cpu(), torch.logsumexp(keys, -1), rtol=1e-5, atol=1e-5)


def test_sampler_philox_replay_and_statistics(drb):
    """In-kernel Philox: dump the noise, replay it through the oracle -> identical samples;
    the noise is Gumbel(0,1) (mean = Euler gamma, var = pi^2/6); different offsets differ."""
    from oracle import sampler as osamp
    B, K, N, s = 2, 50, 2000, 5
    logits = drb.synth.logits_regime(B, N, "L1", seed=3)
    idx, _, _, noise = drb.ops.sample(logits.to(DEV), K, s, 1.0, seed=42, offset=0, want_noise=True)
    for b in range(B):
        _, _, ref_idx = osamp.sample(logits[b], noise[b].cpu(), s, 1.0)
        assert torch.equal(idx[b].cpu().long(), ref_idx)
    n = noise.flatten().double().cpu()
    assert abs(n.mean().item() - 0.5772) < 0.02 and abs(n.var().item() - math.pi ** 2 / 6) < 0.05
    idx2, _, _, _ = drb.ops.sample(logits.to(DEV), K, s, 1.0, seed=42, offset=1)
    assert not torch.equal(idx, idx2)
    idx3, _, _, _ = drb.ops.sample(logits.to(DEV), K, s, 1.0, seed=42, offset=0)
    assert torch.equal(idx, idx3)                                             # deterministic


def test_sampler_race_equals_exact(drb):
    """Test mode uses the exponential-race fast path (one SFU op per element); it must pick the same
    points as the exact key = (logit + G) / tau path that the oracle replays, on the same Philox stream."""
    # the last three rows: N beyond one shared-memory chunk of the training kernel (tables swept in pieces), and
    # peaked weights, where its candidate pre-filter comes up short and the unfiltered second sweep takes over
    for s, N, K, regime in ((5, 2000, 1000, "L0"), (8, 2000, 512, "L1"), (3, 4100, 64, "L1"), (7, 333, 100, "L0"),
                            (3, 12000, 64, "L1"), (5, 50000, 24, "L0"), (8, 2000, 256, "peaked")):
        B = 4
        if regime == "peaked":
            logits = drb.synth.logits_regime(B, N, "L1", seed=s)
            logits[:, ::400] += 9.0
            logits = logits.to(DEV)
        else:
            logits = drb.synth.logits_regime(B, N, regime, seed=s).to(DEV)
        fast, _, _, _ = drb.ops.sample(logits, K, s, 1.0, seed=11, offset=3)
        # want_noise forces the exact-key kernel (the one the oracle replays)
        exact, lse_x, key_x, noise = drb.ops.sample(logits, K, s, 1.0, seed=11, offset=3, want_lse=True,
                                                    want_noise=True)
        same_rows = (fast == exact).all(dim=-1).float().mean().item()
        assert same_rows >= 0.999, (s, N, K, same_rows)
        assert (fast[..., 1:] > fast[..., :-1]).all()          # ascending, distinct
        assert int(fast.min()) >= 0 and int(fast.max()) < N
        # training fast path (tau = 1, in-kernel noise): same draw, same sets, same normaliser and keys
        tr, lse_t, key_t, _ = drb.ops.sample(logits, K, s, 1.0, seed=11, offset=3, want_lse=True)
        rows = (tr == exact).all(dim=-1)
        assert rows.float().mean().item() >= 0.999
        assert torch.allclose(lse_t, lse_x, rtol=2e-5, atol=2e-5)
        assert torch.allclose(key_t[rows], key_x[rows], rtol=1e-5, atol=1e-5)
        # ... and the backward's fast path (noise regenerated) against the generic one (noise re-read)
        g_sel = torch.randn(B, K, s, device=DEV)
        g_fast = drb.ops.sample_backward(logits, exact, lse_x, key_x, g_sel, 1.0, None, 11, 3)
        g_gen = drb.ops.sample_backward(logits, exact, lse_x, key_x, g_sel, 1.0, noise, 11, 3)
        assert (g_fast - g_gen).norm() <= 1e-4 * g_gen.norm()


def test_set_sampler_same_law_as_gumbel_topk(drb):
    """drb_sample_sets draws without replacement by inverse CDF (Plackett-Luce); Gumbel top-s is the same
    law.  Compare, on a small alphabet, (i) the first-pick marginal with softmax(logits), (ii) inclusion
    frequencies and pair co-occurrence frequencies of both kernels over 200 000 hypotheses."""
    B, N, K, s = 1, 12, 200000, 3
    logits = torch.tensor([[0.3, -1.0, 2.0, 0.0, 1.2, -0.5, 0.7, -2.0, 1.5, 0.1, -0.2, 0.9]])
    a = drb.ops.sample_sets(logits.to(DEV), K, s, seed=3, offset=0)[0].cpu().long()
    b = drb.ops.sample(logits.to(DEV), K, s, 1.0, seed=4, offset=0, want_lse=True)[0][0].cpu().long()
    assert (a[:, 1:] > a[:, :-1]).all() and int(a.min()) >= 0 and int(a.max()) < N

    def stats(idx):
        inc = torch.zeros(N)
        pair = torch.zeros(N, N)
        onehot = torch.zeros(idx.shape[0], N).scatter_(1, idx, 1.0)
        inc = onehot.mean(0)
        pair = (onehot.T @ onehot) / idx.shape[0]
        return inc, pair

    ia, pa = stats(a)
    ib, pb = stats(b)
    sigma = (0.25 / K) ** 0.5
    assert (ia - ib).abs().max() < 6 * sigma * 1.5
    assert (pa - pb).abs().max() < 6 * sigma * 1.5
    # exact inclusion probability of the most likely item under Plackett-Luce (enumeration over ordered triples)
    w = torch.softmax(logits[0].double(), 0)
    p_inc = torch.zeros(N, dtype=torch.float64)
    for i in range(N):
        for j in range(N):
            if j == i:
                continue
            for l in range(N):
                if l in (i, j):
                    continue
                p = w[i] * w[j] / (1 - w[i]) * w[l] / (1 - w[i] - w[j])
                p_inc[i] += p
                p_inc[j] += p
                p_inc[l] += p
    assert (ia.double() - p_inc).abs().max() < 8 * sigma
    # determinism / offsets / degenerate weights
    assert torch.equal(drb.ops.sample_sets(logits.to(DEV), 1000, s, seed=3, offset=0),
                       drb.ops.sample_sets(logits.to(DEV), 1000, s, seed=3, offset=0))
    assert not torch.equal(drb.ops.sample_sets(logits.to(DEV), 1000, s, seed=3, offset=0),
                           drb.ops.sample_sets(logits.to(DEV), 1000, s, seed=3, offset=1))
    spike = torch.full((2, 500), -80.0)
    spike[:, 17] = 0.0
    d = drb.ops.sample_sets(spike.to(DEV), 64, 5, seed=1, offset=0).cpu()
    assert (d[..., 1:] > d[..., :-1]).all() and (d == 17).any(dim=-1).all()
    big = drb.ops.sample_sets(torch.zeros(2, 50000, device=DEV), 128, 3, seed=1, offset=0)     # cfg4 size
    assert int(big.max()) < 50000 and (big[..., 1:] > big[..., :-1]).all()


def test_sampler_follows_the_weights(drb):
    """Gumbel-max sampling draws the first pick with probability softmax(logits): empirical
    inclusion frequencies over many hypotheses follow the weights (fast path)."""
    B, N, K, s = 1, 64, 20000, 3
    logits = torch.linspace(-2.0, 2.0, N)[None]
    idx, _, _, _ = drb.ops.sample(logits.to(DEV), K, s, 1.0, seed=5, offset=0)
    counts = torch.bincount(idx.flatten().cpu().long(), minlength=N).double()
    # inclusion probability is monotone in the logit; compare halves and the rank correlation
    assert counts[N // 2:].sum() > 2.5 * counts[: N // 2].sum()
    r = torch.corrcoef(torch.stack((counts, torch.softmax(logits[0].double(), 0))))[0, 1]
    assert r > 0.98


# ---- a3 five-point ---------------------------------------------------------------------------------
def _e5_match_stats(models, nsol, ref64):
    K = ref64.shape[0] // 10
    ref = ref64.view(K, 10, 3, 3)
    real = (trace_constraint_residual(ref64) < 1e-8).view(K, 10)
    d = match_up_to_sign(models.view(K, 10, 3, 3).cpu(), ref)[real]
    return d, real


def test_e5_models_vs_fp64_reference(drb, golden):
    g = golden("nister")
    models, nsol = drb.ops.solve_e5(g["pts"].to(DEV))
    d, real = _e5_match_stats(models[0], nsol[0], g["E64"])
    d_ref32, _ = _e5_match_stats(g["E32"], None, g["E64"])
    # every genuine (real-root) model of the fp64 reference must be found; the fp32 reference itself
    # only reproduces ~82 % of them at 1e-3 (its own noise floor, DESIGN.md "Parity contract")
    assert (d < 1e-3).float().mean() >= 0.95
    assert (d < 1e-3).float().mean() >= (d_ref32 < 1e-3).float().mean()
    assert d.median() < 1e-5
    # emitted models are genuine essential matrices, unit norm, fitting their sample
    valid = torch.arange(10)[None] < nsol[0].cpu()[:, None]
    m = models[0].cpu()[valid]
    assert (trace_constraint_residual(m) < 1e-3).float().mean() > 0.97
    assert torch.allclose(m.flatten(1).norm(dim=1), torch.ones(m.shape[0]), atol=1e-5)
    pts = g["pts"]
    h1 = torch.cat((pts[..., :2], torch.ones_like(pts[..., :1])), -1).double()
    h2 = torch.cat((pts[..., 2:], torch.ones_like(pts[..., :1])), -1).double()
    r = torch.einsum("kni,ksij,knj->ksn", h2, models[0].cpu().double(), h1)
    assert r[valid].abs().max() < 1e-4
    # unused slots are the identity, like the reference's padding (nister.py:400-401)
    assert torch.equal(models[0].cpu()[~valid], torch.eye(3).expand((~valid).sum(), 3, 3))


def test_e5_indexed_equals_gathered(drb, golden):
    g = golden("nister")
    m_g, n_g = drb.ops.solve_e5(g["pts"].to(DEV))
    m_i, n_i = drb.ops.solve_e5(g["matches"][None].to(DEV), g["idx"][None].to(DEV))
    assert torch.equal(m_g, m_i) and torch.equal(n_g, n_i)


def test_stewenius_solutions_contained(drb, golden):
    """The Stewenius class solves the same system: every real-eigenvalue model of the reference
    (up to scale and sign) is among ours."""
    g = golden("stewenius")
    models, nsol = drb.ops.solve_e5(g["pts"].to(DEV))
    ref = unit(g["E32"]).view(-1, 10, 3, 3)
    real = (trace_constraint_residual(unit(g["E32"])) < 1e-4).view(-1, 10)
    d = match_up_to_sign(models[0].cpu(), ref)[real]
    assert (d < 2e-3).float().mean() > 0.9


def test_stewenius_fp64_solution_set_is_found(drb, golden):
    """Row a4: why one kernel serves both of the reference's five-point classes.  Run in fp64 (stewenius_64.npz, the
    reference class itself; tests/test_oracle_golden.py shows its genuine models coincide with the fp64 Nister class's
    to 1e-8, 224 of 224), the action-matrix solver (stewenius.py:20-80) defines a SET of essential matrices per
    sample; the kernel finds that set -- the same bar as for the Nister class (>= 95 % within 1e-3 up to sign and
    scale, median < 1e-5; measured 98.7 %, median 8e-7).  For comparison the reference's own fp32 run of this class
    reaches 99.6 % of its fp64 models at 1e-3 with a median of 4e-6 (LAPACK's eig is more robust than Sturm + polish
    on the few near-double roots, less precise on the rest); its fp32 Nister class reaches 82 %."""
    g, g64 = golden("stewenius"), golden("stewenius_64")
    K = g["pts"].shape[0]
    ref = unit(g64["E64"]).view(K, 10, 3, 3)
    genuine = (trace_constraint_residual(unit(g64["E64"])) < 1e-8).view(K, 10)
    models, nsol = drb.ops.solve_e5(g["pts"].to(DEV))
    ours = models[0].cpu()
    live = torch.arange(10)[None] < nsol[0].cpu()[:, None]
    ours = torch.where(live[..., None, None], ours, torch.full_like(ours, 1e3))
    d = match_up_to_sign(ours, ref)[genuine]
    assert (d < 1e-3).float().mean() >= 0.95 and d.median() < 1e-5
    # the reference's own fp32 run of the class against its fp64 run, same measure: more models within 1e-3 (eig
    # copes better with the near-double roots), but a 5x larger typical error
    ref32 = unit(g["E32"]).view(K, 10, 3, 3)
    d32 = match_up_to_sign(ref32, ref)[genuine]
    assert d.median() < d32.median()
    # and the host mirror class routes to the same kernel with the reference's call shape
    from differentiable_ransac_b200.estimators.essential_matrix_estimator_stewenius import EssentialMatrixEstimator
    est = EssentialMatrixEstimator(DEV)
    out = est.estimate_minimal_model(g["pts"].to(DEV))
    assert out.shape == (K * 10, 3, 3) and torch.equal(out.view(K, 10, 3, 3), models[0])


# ---- a9 MSAC + arg-max -----------------------------------------------------------------------------------
def test_msac_scores_argmax_mask(drb, golden):
    g = golden("msac")
    thr = torch.tensor([float(g["threshold"])])
    scores, best = drb.ops.score_msac(g["matches"][None].to(DEV), g["models"][None].to(DEV), thr.to(DEV))
    assert torch.allclose(scores[0].cpu(), g["scores"], rtol=1e-4, atol=1e-4)
    bid, bscore, bmodel, mask, ninl = drb.ops.best_finalize(g["matches"][None].to(DEV), g["models"][None].to(DEV),
                                                            best, thr.to(DEV))
    assert int(bid[0]) == int(g["best"])
    assert torch.equal(bmodel[0].cpu(), g["models"][int(g["best"])])
    assert (mask[0].cpu().bool() != g["best_mask"]).sum() <= 1 and abs(int(ninl[0]) - int(g["best_mask"].sum())) <= 1
    assert abs(float(bscore[0]) - float(g["scores"].max())) < 1e-4 * float(g["scores"].max())


def test_msac_ragged_counts_and_ids(drb):
    """count / ids (the compact list produced by the solver), N not a multiple of the tile,
    an empty pair and B > 1."""
    from oracle import scoring
    B, M, N = 3, 300, 2500
    matches, _, _ = drb.synth.relative_pose_batch(B, N, seed=5)
    gen = torch.Generator().manual_seed(0)
    models = unit(torch.randn(B, M, 3, 3, generator=gen))
    count = torch.tensor([300, 0, 129], dtype=torch.int32)
    ids = torch.stack([torch.randperm(1000, generator=gen)[:M] for _ in range(B)]).int()
    thr = torch.tensor([0.01, 0.02, 0.005])
    scores, best = drb.ops.score_msac(matches.to(DEV), models.to(DEV), thr.to(DEV), count=count.to(DEV),
                                      ids=ids.to(DEV))
    best = best.cpu()
    for b in range(B):
        c = int(count[b])
        if c == 0:
            assert int(best[b]) == 0
            continue
        ref, _ = scoring.msac_score(matches[b], models[b, :c], float(thr[b]))
        assert torch.allclose(scores[b, :c].cpu(), ref, rtol=1e-4, atol=1e-4)
        key = int(best[b]) & 0xFFFFFFFFFFFFFFFF
        assert 0xFFFFFFFF - (key & 0xFFFFFFFF) == int(ids[b, int(torch.argmax(ref))])


# ---- a11 driver test mode -----------------------------------------------------------------------------------
def test_pipeline_matches_reference_driver(drb, golden):
    """Same noise as the reference loop body (two chunks of 32 = one batch of 64 here): identical
    minimal samples, identical best hypothesis, winner model up to sign, mask IoU."""
    g = golden("driver_test")
    noise = g["noise"].reshape(1, 64, -1)
    thr = torch.tensor([float(g["threshold"])])
    out = drb.engine.ransac_e5_test(g["matches"][None].to(DEV), g["logits"][None].to(DEV), 64, thr.to(DEV),
                                    noise=noise.to(DEV), want_scores=True)
    assert torch.equal(out["idx"][0].cpu().long(), g["idx"].reshape(64, 5))
    ref_hyp = int(g["best_chunk"]) * 32 + int(g["best_hyp"])
    assert int(out["best_hyp"][0]) == ref_hyp
    bm, rm = out["best_model"][0].cpu(), g["best_model"]
    assert min((bm - rm).norm(), (bm + rm).norm()) < 2e-3
    inter = (out["mask"][0].cpu() & g["best_mask"]).sum().item()
    union = (out["mask"][0].cpu() | g["best_mask"]).sum().item()
    assert inter / max(union, 1) > 0.95


# ---- a5 / a6 fundamental ----------------------------------------------------------------------------------------
def test_f8_vs_reference(drb, golden):
    g = golden("f8")
    F, valid = drb.ops.solve_f8(g["pts"].to(DEV))
    assert valid.all()
    F = F[0].cpu()
    ref = g["F64"].float()
    d = torch.minimum((F - ref).flatten(1).norm(dim=1), (F + ref).flatten(1).norm(dim=1)) / ref.flatten(1).norm(dim=1)
    assert d.max() < 1e-4         # same scale (unit-norm null vector, de-normalised), sign free
    F2, _ = drb.ops.solve_f8(g["matches"][None].to(DEV), g["idx"][None].to(DEV))
    assert torch.equal(F2[0].cpu(), F)


def test_f7_algebra(drb, golden):
    from oracle import fundamental
    g = golden("f8")
    pts = g["pts"][:, :7].contiguous()
    F, nsol = drb.ops.solve_f7(pts.to(DEV))
    Fo, valid = fundamental.seven_point(pts.double())
    F, nsol = F[0].cpu().double(), nsol[0].cpu()
    assert (nsol == valid.sum(1)).float().mean() > 0.95
    h1 = torch.cat((pts[..., :2], torch.ones_like(pts[..., :1])), -1).double()
    h2 = torch.cat((pts[..., 2:], torch.ones_like(pts[..., :1])), -1).double()
    # pixel coordinates ~ 1e2..1e3: scale the residual by the magnitude of the terms
    scale = torch.einsum("kni,ksij,knj->ksn", h2.abs(), F.abs(), h1.abs())
    r = torch.einsum("kni,ksij,knj->ksn", h2, F, h1) / scale
    ok = torch.arange(3)[None] < nsol[:, None]
    assert r[ok].abs().max() < 1e-4
    assert torch.linalg.det(unit(F[ok])).abs().max() < 1e-4


# ---- a7 / a8 rigid ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flag", [1, 0])
def test_rigid_solver_and_residual(drb, golden, flag):
    g = golden("rigid")
    m, valid = drb.ops.solve_rigid3(g["pts"].to(DEV), flag=bool(flag))
    assert valid.all()
    assert torch.allclose(m[0].cpu(), g[f"model_{flag}"], atol=5e-4, rtol=1e-4)
    res, ninl = drb.ops.rigid_residual_forward(g["points"][None].to(DEV), g[f"model_{flag}"][None].to(DEV))
    assert torch.allclose(res[0].cpu(), g[f"res_{flag}"], rtol=1e-4)
    assert (ninl[0].cpu().long() - g[f"ninl_{flag}"]).abs().max() <= 1


# ---- a10 symmetric epipolar loss ------------------------------------------------------------------------------------
def test_episym_forward_backward(drb, golden):
    from oracle import scoring
    g = golden("episym")
    pts = g["matches"][g["gt_mask"]]
    P = pts.shape[0]
    row = drb.ops.episym_forward(pts[None].to(DEV), g["models"][None].to(DEV))
    assert torch.allclose(row[0].cpu() / P, g["row_mean"], rtol=1e-4, atol=1e-6)
    # padded + npts path gives the same sums
    padded = torch.cat((pts, torch.zeros(37, 4)))
    row2 = drb.ops.episym_forward(padded[None].to(DEV), g["models"][None].to(DEV),
                                  npts=torch.tensor([P], dtype=torch.int32, device=DEV))
    assert torch.allclose(row2, row, rtol=1e-5)
    # backward against autograd of the oracle (fp64)
    md = g["models"][:64].double().clone().requires_grad_(True)
    K = md.shape[0]
    e = scoring.episym(pts[:, :2].double().repeat(K, 1, 1), pts[:, 2:].double().repeat(K, 1, 1), md)
    gr = torch.randn(K, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    (torch.min(e, torch.ones_like(e)).sum(1) * gr).sum().backward()
    gm = drb.ops.episym_backward(pts[None].to(DEV), g["models"][None, :64].to(DEV), gr.float()[None].to(DEV))
    rel = (gm[0].cpu().double() - md.grad).flatten(1).norm(dim=1) / md.grad.flatten(1).norm(dim=1)
    assert rel.max() < 1e-3 and rel.median() < 1e-5


# ---- backward of the whole train path ---------------------------------------------------------------------------------
def test_e5_train_step_gradients_vs_reference(drb, golden):
    """cfg5 unit: sample -> 5pt -> closest-to-GT -> clamped episym mean; d loss / d logits and
    d loss / d matches against the reference's own autograd in fp64."""
    g64, g32 = golden("driver_train_64"), golden("driver_train_32")
    matches = g64["matches"][None].to(DEV).requires_grad_(True)
    logits = g64["logits"][None].to(DEV).requires_grad_(True)
    noise = g64["noise"].reshape(1, 64, -1).to(DEV)
    # reference_sign: the reference picks the slot closest to GT without sign handling
    chosen, valid = drb.engine.HypothesizeE5.apply(matches, logits, g64["E_gt"][None].to(DEV), 64, 1.0, noise, 0, 0,
                                                   True)
    inl = g64["matches"][g64["gt_mask"]]
    loss = drb.engine.match_loss(chosen, valid, inl[None].to(DEV))[0]
    loss.backward()
    # The reference driver's own slot choice is sign-noise driven, so the comparison is
    # with the `sel_*` golden: same reference sampler/estimator/loss/autograd, slot = genuine model
    # closest to GT up to sign (tests/golden/make_golden.py).
    assert torch.equal(valid[0].cpu(), g64["sel_keep"])
    ref_models = g64["sel_models"]
    d = torch.minimum((chosen[0].detach().cpu() - ref_models).flatten(1).norm(dim=1),
                      (chosen[0].detach().cpu() + ref_models).flatten(1).norm(dim=1))
    same = d < 1e-3
    assert same.float().mean() >= 0.95
    gl, rl, rl32 = logits.grad[0].cpu().double(), g64["sel_grad_logits"], g32["sel_grad_logits"].double()
    gm, rm, rm32 = matches.grad[0].cpu().double(), g64["sel_grad_matches"], g32["sel_grad_matches"].double()
    err_l, err_m = (gl - rl).norm() / rl.norm(), (gm - rm).norm() / rm.norm()
    # identical model set (measured on the B200: every model within 1e-3, gradients 1.5e-5 / 2.3e-5 from the fp64
    # reference, profiles/r2_grad_parity.json): soft loss and gradients within 1e-4 relative of the fp64 reference
    assert same.all()
    assert abs(loss.item() - g64["sel_loss"].item()) < 1e-4 * abs(g64["sel_loss"].item())
    assert err_l < 1e-4 and err_m < 1e-4
    # in any case: at least as close to the fp64 reference as the fp32 reference itself is
    assert err_l < max(1e-4, 1.5 * (rl32 - rl).norm() / rl.norm())
    assert err_m < max(1e-4, 1.5 * (rm32 - rm).norm() / rm.norm())


def test_f8_train_step_gradients_vs_reference(drb, golden):
    """The eight-point unit in PIXEL units.  This fixture's gradient is ill-conditioned by construction: 98.7 % of its
    (model, point) terms sit above the clamp at 1, and merely rounding the reference's exact fp64 models to fp32
    moves d loss / d models by 2.2 % (the fp32 reference's own gradient is 188 % away from its fp64 self).  Measured
    on the B200: 9.2e-3 / 7.3e-3 (profiles/r2_grad_parity.json) -- hence the 1e-2 bar here; the well-conditioned
    chain train.py actually runs is `test_f8_layer_train_gradients_vs_reference` below, at 1e-3."""
    g64 = golden("f8_train_64")
    matches = g64["matches"][None].to(DEV).requires_grad_(True)
    logits = g64["logits"][None].to(DEV).requires_grad_(True)
    noise = g64["noise"][None].to(DEV)
    models, valid = drb.engine.HypothesizeF8.apply(matches, logits, 48, 1.0, noise, 0, 0)
    assert valid.all()
    ref = g64["models"]
    d = torch.minimum((models[0].detach().cpu() - ref).flatten(1).norm(dim=1),
                      (models[0].detach().cpu() + ref).flatten(1).norm(dim=1)) / ref.flatten(1).norm(dim=1)
    assert d.max() < 1e-3
    inl = g64["matches"][g64["gt_mask"]]
    loss = drb.engine.match_loss(models, valid, inl[None].to(DEV))[0]
    loss.backward()
    assert abs(loss.item() - g64["loss"].item()) < 1e-3 * abs(g64["loss"].item())
    gl, rl = logits.grad[0].cpu().double(), g64["grad_logits"]
    assert (gl - rl).norm() / rl.norm() < 1e-2
    gm, rm = matches.grad[0].cpu().double(), g64["grad_matches"]
    assert (gm - rm).norm() / rm.norm() < 1e-2


def test_f8_layer_train_gradients_vs_reference(drb, golden):
    """cfg3 as train.py runs it (`-fmat 1 -sam 3 -tr 1 -w2 1`): RANSACLayer.forward (points denormalised to pixels,
    eight-point per sample) -> MatchLoss(fmat=1) (E = K2^T F K1 on K-normalised points) -> backward to the sampling
    weights, against the reference's own run of the same chain in fp64 (tests/golden/make_golden_r2.py)."""
    import types

    from differentiable_ransac_b200.loss import MatchLoss
    from differentiable_ransac_b200.model_cl import RANSACLayer
    g64, g32 = golden("f8_layer_train_64"), golden("f8_layer_train_32")
    K = g64["noise"].shape[0]
    opt = types.SimpleNamespace(device=DEV, fmat=1, sampler=3, precision=1, tr=1, threshold=0.75, ransac_batch_size=K,
                                weighted=0)
    layer = RANSACLayer(opt)
    layer.estimator.max_iterations = K
    layer.estimator.sampler.injected_noise = g64["noise"].to(DEV)
    pts = g64["points"].to(DEV)
    w = g64["logits"].to(DEV).requires_grad_(True)
    Kc, im = g64["K"].to(DEV), g64["im_size"].to(DEV)
    Es, _ = layer.forward(pts, w, Kc, Kc, im, im, g64["F_gt"].to(DEV))
    loss = MatchLoss(1).forward([Es], g64["gt_E"].numpy()[None], [pts[:, 0:2]], [pts[:, 2:4]], [Kc], [Kc], [im], [im])
    loss.backward()
    ref = g64["models"]
    d = torch.minimum((Es.detach().cpu() - ref).flatten(1).norm(dim=1),
                      (Es.detach().cpu() + ref).flatten(1).norm(dim=1)) / ref.flatten(1).norm(dim=1)
    assert Es.shape == ref.shape and d.max() < 1e-3
    assert abs(loss.item() - g64["loss"].item()) < 1e-4 * abs(g64["loss"].item())
    gl, rl, rl32 = w.grad.cpu().double(), g64["grad_logits"], g32["grad_logits"].double()
    err = (gl - rl).norm() / rl.norm()
    assert err < 1e-3, float(err)
    assert err < 1.5 * (rl32 - rl).norm() / rl.norm()       # at least as close to fp64 as the fp32 reference is (3.1e-4)


def test_rigid_train_step_vs_reference(drb, golden):
    g = golden("rigid_train")
    points = g["points"][None].to(DEV)
    logits = g["logits"][None].to(DEV).requires_grad_(True)
    noise = g["noise"].reshape(1, 64, -1).to(DEV)
    models, valid = drb.engine.HypothesizeRigid.apply(points, logits, 64, True, 1.0, noise, 0, 0)
    assert valid.all()
    # flag=True: the reference's R is I + fp32 SVD noise, times centroids of size ~10
    assert torch.allclose(models[0].detach().cpu(), g["models"], atol=5e-3, rtol=1e-4)
    res = drb.engine.RigidResidual.apply(points, models)
    assert torch.allclose(res[0].detach().cpu(), g["residuals"], rtol=2e-3)
    loss = res.mean()
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) < 2e-3 * abs(g["loss"].item())
    gl, rl = logits.grad[0].cpu(), g["grad_logits"]
    assert (gl - rl).norm() / rl.norm() < 2e-4           # the fp32 reference: itself 9.2e-5 from its fp64 run
    # against the reference's fp64 run of the same branch (rigid_train_64): measured 4.8e-7 on the B200
    g64 = golden("rigid_train_64")
    assert abs(loss.item() - g64["loss"].item()) < 1e-5 * abs(g64["loss"].item())
    assert torch.allclose(res[0].detach().cpu().double(), g64["residuals"], rtol=1e-5)
    assert (gl.double() - g64["grad_logits"]).norm() / g64["grad_logits"].norm() < 1e-4


# ---- full BASELINE sizes through size-independent properties -----------------------------------------------------------
def test_headline_shape_properties(drb):
    """cfg2 shape (B=32, K=1000, N=2000): determinism, winner = arg-max of the dense scores, the
    winner's inlier count recomputed by the oracle, and the GT pose is recovered on clean pairs."""
    from oracle import scoring
    B, K, N = 32, 1000, 2000
    matches, E_gt, inl = drb.synth.relative_pose_batch(B, N, seed=1234)
    logits = drb.synth.logits_regime(B, N, "L0", seed=99)
    thr = torch.full((B,), 0.75 / 800.0)
    a = drb.engine.ransac_e5_test(matches.to(DEV), logits.to(DEV), K, thr.to(DEV), seed=42, want_scores=True)
    b = drb.engine.ransac_e5_test(matches.to(DEV), logits.to(DEV), K, thr.to(DEV), seed=42)
    assert torch.equal(a["best_id"], b["best_id"]) and torch.equal(a["best_model"], b["best_model"])
    cc = a["ccount"].cpu()
    for bi in range(B):
        sc = a["scores"][bi, : cc[bi]].cpu()
        ids = a["cids"][bi, : cc[bi]].cpu()
        top = sc.max()
        assert float(a["best_score"][bi]) == float(top)
        assert int(a["best_id"][bi]) == int(ids[sc == top].min())
    # oracle re-scores the winners
    for bi in (0, 1, 2, 17):
        s, m = scoring.msac_score(matches[bi], a["best_model"][bi].cpu()[None], float(thr[bi]))
        assert abs(s[0].item() - a["best_score"][bi].item()) < 1e-3 * max(1.0, s[0].item())
        assert abs(int(m.sum()) - int(a["ninl"][bi])) <= 2
    # noise-free synthetic pairs: the winner is the true essential matrix wherever 1000 draws contain an
    # all-inlier sample with near certainty (inlier ratio >= 0.4; at 0.2, 0.2^5 * 1000 = 0.3 such samples)
    bm = a["best_model"].cpu()
    err = torch.minimum((bm - E_gt).flatten(1).norm(dim=1), (bm + E_gt).flatten(1).norm(dim=1))
    easy = torch.arange(B) % 3 != 0
    assert (err[easy] < 1e-2).float().mean() > 0.9
