"""GPU parity of what follows the hypothesize-and-score loop in test mode: the
non-minimal fits (`drb_refit_e5`, `drb_refit_f8`), local optimisation and the final refit, against
fixtures produced by the reference itself (tests/golden/make_golden.py: refit_*, driver_full_*)."""
import types

import pytest
import torch

from helpers import match_up_to_sign, trace_constraint_residual, unit

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_refit_e5_finds_every_genuine_reference_model(golden):
    from differentiable_ransac_b200 import ops
    g = golden("refit_e5")
    m = g["matches"].to(DEV)[None]
    for ref, mask in ((g["E_all64"], None), (g["E_inl64"], g["mask"].bool().to(DEV)[None])):
        models, nsol = ops.refit_e5(m, mask)
        n = int(nsol[0])
        genuine = trace_constraint_residual(ref) < 1e-8          # the rest are complex-root leftovers (D3)
        assert n == int(genuine.sum()) > 0
        ours = models[0, :n].cpu()
        assert trace_constraint_residual(ours).max() < 1e-5
        assert (ours.flatten(1).norm(dim=1) - 1).abs().max() < 1e-6
        assert match_up_to_sign(ours[None], unit(ref[genuine])[None])[0].max() < 1e-6
        eye = torch.eye(3)[None].expand(10 - n, 3, 3)
        assert torch.equal(models[0, n:].cpu(), eye)             # identity padding, nister.py:400-401


def test_refit_f8_matches_reference_plain_and_weighted(golden):
    from differentiable_ransac_b200 import ops
    g = golden("refit_f8")
    m, mask = g["matches"].to(DEV)[None], g["mask"].bool().to(DEV)[None]
    for ref, w in ((g["F_inl64"], None), (g["F_w64"], g["weights"].to(DEV)[None])):
        F, ok = ops.refit_f8(m, mask, w)
        assert int(ok[0]) == 1
        Fc = F[0, 0].double().cpu()
        assert min((Fc - ref[0]).abs().max(), (Fc + ref[0]).abs().max()) < 1e-6     # same scale as the reference
    # the fp32 reference is this far from its fp64 self; we are closer to fp64 than it is
    Fc = ops.refit_f8(m, mask)[0][0, 0].double().cpu()
    ours = min((Fc - g["F_inl64"][0]).norm(), (Fc + g["F_inl64"][0]).norm())
    theirs = (g["F_inl32"][0].double() - g["F_inl64"][0]).norm()
    assert ours <= theirs + 1e-9


def test_refit_edge_cases_and_batching(golden):
    from differentiable_ransac_b200 import ops
    g5, g8 = golden("refit_e5"), golden("refit_f8")
    # too few selected correspondences -> no model, identity out
    m = g8["matches"].to(DEV)[None]
    few = torch.zeros(1, 2000, dtype=torch.bool, device=DEV)
    few[0, :7] = True
    F, ok = ops.refit_f8(m, few)
    assert int(ok[0]) == 0 and torch.equal(F[0, 0].cpu(), torch.eye(3))
    E, ns = ops.refit_e5(m, few[:, :2000] & (torch.arange(2000, device=DEV) < 4)[None])
    assert int(ns[0]) == 0 and torch.equal(E[0].cpu(), torch.eye(3)[None].expand(10, 3, 3))
    # mask = None equals an all-ones mask; pairs of a batch are independent
    m5 = g5["matches"].to(DEV)
    batch = torch.stack((m5, m5.flip(0), m5))
    masks = torch.stack((torch.ones(1000, dtype=torch.bool, device=DEV), g5["mask"].bool().to(DEV).flip(0),
                         g5["mask"].bool().to(DEV)))
    Eb, nb = ops.refit_e5(batch, masks)
    E0, n0 = ops.refit_e5(m5[None], None)
    E2, n2 = ops.refit_e5(m5[None], masks[2:3])
    assert torch.equal(Eb[0], E0[0]) and int(nb[0]) == int(n0[0])
    assert torch.equal(Eb[2], E2[0]) and int(nb[2]) == int(n2[0])
    # the flipped pair holds the same correspondences in another order: same models up to summation order
    assert int(nb[1]) == int(nb[2])
    assert match_up_to_sign(Eb[1, :int(nb[1])].cpu()[None], Eb[2, :int(nb[2])].cpu()[None])[0].max() < 1e-6


def test_nonminimal_input_through_the_estimator_plugins(golden):
    """`estimator.estimate_model(points[1,n,4])` with n > sample_size, as ransac.py:151-165, :231-240 call it."""
    from differentiable_ransac_b200.estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
    from differentiable_ransac_b200.estimators.fundamental_matrix_estimator import FundamentalMatrixEstimatorNew
    g5, g8 = golden("refit_e5"), golden("refit_f8")
    est = EssentialMatrixEstimatorNister(DEV)
    E = est.estimate_model(g5["matches"][g5["mask"].bool()].to(DEV)[None])
    assert E.shape == (10, 3, 3)
    n = int(est.last_nsol[0])
    genuine = trace_constraint_residual(g5["E_inl64"]) < 1e-8
    assert match_up_to_sign(E[:n].cpu()[None], unit(g5["E_inl64"][genuine])[None])[0].max() < 1e-6
    F = FundamentalMatrixEstimatorNew(DEV).estimate_model(g8["matches"][g8["mask"].bool()].to(DEV)[None])
    assert F.shape == (1, 3, 3)
    assert min((F[0].cpu().double() - g8["F_inl64"][0]).abs().max(),
               (F[0].cpu().double() + g8["F_inl64"][0]).abs().max()) < 1e-6


def _driver(fmat, lo):
    from differentiable_ransac_b200.model_cl import RANSACLayer
    opt = types.SimpleNamespace(device=DEV, fmat=int(fmat), sampler=3 if fmat else 2, precision=1, tr=0, threshold=0.75,
                                ransac_batch_size=32, weighted=0)
    drv = RANSACLayer(opt).estimator
    drv.max_iterations, drv.lo, drv.lo_iters = 128, lo, 8
    return drv


@pytest.mark.parametrize("name,fmat", [("driver_full_e5_lo0", False), ("driver_full_e5_lo2", False),
                                       ("driver_full_f8_lo0", True), ("driver_full_f8_lo2", True)])
def test_full_test_mode_driver_vs_reference(golden, name, fmat):
    """`RANSAC.__call__` in test mode with the reference's injected noise: chunked loop, adaptive exit, LO on every
    improvement (lo=2), final refit -- the reference's own end-to-end answer.  For the five-point path the bar is
    the reference run in fp64 (`*_64`): its fp32 run loses genuine models to LAPACK rounding (here the
    winner of chunk 1), so it is only required that we are at least as close to fp64 as the fp32 reference is."""
    from differentiable_ransac_b200 import synth
    g = golden(name if fmat else name + "_64")
    m = g["matches"]
    Kc = g["K"] if fmat else g["K1"]
    drv = _driver(fmat, int(name[-1]))
    noise = torch.cat([synth.gumbel_noise((32, m.shape[0]), seed=int(sd)) for sd in g["noise_seeds"]])
    drv.sampler.injected_noise = noise.to(DEV)
    model, mask, score, its = drv(m.to(DEV), g["logits"].to(DEV), Kc, Kc, None)
    assert its == int(g["iterations"])
    ref_model, ref_mask = g["best_model"].float(), g["best_mask"].bool()
    rel = abs(float(score) - float(g["best_score"])) / float(g["best_score"])
    mu, ru = unit(model.cpu()), unit(ref_model)
    dist = float(min((mu - ru).norm(), (mu + ru).norm()))
    iou = (mask.cpu() & ref_mask).sum().item() / max((mask.cpu() | ref_mask).sum().item(), 1)
    assert rel < 2e-3 and dist < 2e-3 and iou > 0.98, (rel, dist, iou)
    if not fmat:
        g32 = golden(name)
        assert abs(float(score) - float(g["best_score"])) <= abs(float(g32["best_score"]) - float(g["best_score"])) + 0.5


def test_batched_refit_improves_or_keeps_every_pair():
    """`RANSAC.batched_test` (B pairs at once) with LO + final refit: scores never drop, poses stay right."""
    from differentiable_ransac_b200 import synth
    B, N, K = 6, 2000, 512
    pairs = [synth.relative_pose_pair(N, (0.5, 0.6)[b % 2], seed=70 + b, noise=3e-4) for b in range(B)]
    m, E_gt = torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])
    lg = synth.logits_regime(B, N, "L0", seed=3)
    thr = torch.full((B,), 0.75 / 800.0)
    drv = _driver(False, 0)
    drv.final_refit = False
    drv.sampler.seed = 11
    base = drv.batched_test(m.to(DEV), lg.to(DEV), thr.to(DEV), K=K)
    drv2 = _driver(False, 2)
    drv2.sampler.seed = 11
    out = drv2.batched_test(m.to(DEV), lg.to(DEV), thr.to(DEV), K=K)
    assert (out["best_score"] >= base["best_score"] - 1e-3).all()
    assert (out["best_score"] > base["best_score"] * 1.01).any()            # LO on noisy data does help somewhere
    for b in range(B):
        Eu, Gu = unit(out["best_model"][b].cpu()), unit(E_gt[b])
        assert min((Eu - Gu).norm(), (Eu + Gu).norm()) < 5e-2
