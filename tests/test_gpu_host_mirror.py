"""GPU tests of the host mirror: the reference's plugin / driver / layer classes re-provided over
the CUDA path, exercised the way the reference's scripts call them."""
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


def _opt(**kw):
    d = dict(device=DEV, fmat=0, sampler=2, precision=1, tr=0, threshold=0.75, ransac_batch_size=32, weighted=0)
    d.update(kw)
    return types.SimpleNamespace(**d)


K1 = torch.tensor([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])


def test_sampler_dense_return_matches_reference(golden):
    from differentiable_ransac_b200.samplers.gumbel_sampler import GumbelSoftmaxSampler
    g = golden("sampler_L0")
    smp = GumbelSoftmaxSampler(12, 5, device=DEV)
    smp.injected_noise = g["noise"].to(DEV)
    lg = g["logits"].to(DEV).requires_grad_(True)
    ret, y_soft = smp.sample(lg)
    assert torch.equal((ret != 0).nonzero()[:, 1].view(12, -1).cpu(), g["idx"])
    assert torch.allclose(y_soft.cpu(), g["y_soft"], rtol=1e-5, atol=1e-9)
    assert torch.allclose(ret.detach().cpu(), g["ret"], atol=1e-6)
    # the reference's own gather on our dense return (ransac.py:64-65)
    m = g["matches"].to(DEV)
    pts = m.repeat([12, 1, 1]) * ret.unsqueeze(-1)
    assert torch.equal(pts[ret != 0].view(12, -1, 4).cpu(), g["minimal"])
    (ret * torch.randn_like(ret)).sum().backward()
    assert torch.isfinite(lg.grad).all() and lg.grad.abs().sum() > 0


def test_msac_plugin(golden):
    from differentiable_ransac_b200.scorings.msac_score import MSACScore
    g = golden("msac")
    scores, masks = MSACScore(DEV).score(g["matches"].to(DEV), g["models"].to(DEV), float(g["threshold"]))
    assert torch.allclose(scores.cpu(), g["scores"], rtol=1e-4, atol=1e-4)
    bi = torch.argmax(scores)
    assert int(bi) == int(g["best"])
    assert (masks[bi].cpu() != g["best_mask"]).sum() <= 1


def test_estimator_plugins_shapes_and_autograd(golden):
    from differentiable_ransac_b200.estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
    from differentiable_ransac_b200.estimators.fundamental_matrix_estimator import FundamentalMatrixEstimatorNew
    from differentiable_ransac_b200.estimators.rigid_transformation_SVD_based_solver import RigidTransformationSVDBasedSolver
    from oracle import nister
    g = golden("nister")
    est = EssentialMatrixEstimatorNister(DEV)
    pts = g["pts"][:16].to(DEV).requires_grad_(True)
    E = est.estimate_model(pts)
    assert E.shape == (160, 3, 3)
    # gradient of <G, E_slot> against the oracle's autograd (fp64) on matched genuine slots
    p64 = g["pts"][:16].double().clone().requires_grad_(True)
    Eo = nister.five_point(p64).view(16, 10, 3, 3)
    Ev = E.view(16, 10, 3, 3)
    real = (trace_constraint_residual(Eo.detach().reshape(-1, 3, 3)) < 1e-9).view(16, 10)
    G = torch.zeros(16, 10, 3, 3)
    Go = torch.zeros(16, 10, 3, 3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(0)
    used = 0
    for k in range(16):
        for s in range(int(est.last_nsol[k])):
            d_pos = (Eo[k].detach() - Ev[k, s].detach().cpu().double()).flatten(1).norm(dim=1)
            d_neg = (Eo[k].detach() + Ev[k, s].detach().cpu().double()).flatten(1).norm(dim=1)
            d, j = torch.min(torch.minimum(d_pos, d_neg), 0)
            if d < 1e-4 and real[k, j] and Go[k, j].abs().sum() == 0:
                w = torch.randn(3, 3, generator=gen)
                sgn = 1.0 if d_pos[j] < d_neg[j] else -1.0
                G[k, s] = w
                Go[k, j] = sgn * w.double()
                used += 1
    assert used > 30
    (E.view(16, 10, 3, 3) * G.to(DEV)).sum().backward()
    (Eo * Go).sum().backward()
    rel = (pts.grad.cpu().double() - p64.grad).flatten(1).norm(dim=1) / p64.grad.flatten(1).norm(dim=1).clamp_min(1e-12)
    assert rel.median() < 1e-3 and (rel < 5e-2).float().mean() > 0.8
    f = golden("f8")
    fe = FundamentalMatrixEstimatorNew(DEV)
    assert fe.estimate_model(f["pts"].to(DEV)).shape == (64, 3, 3)
    assert fe.estimate_model(f["pts"][:, :7].contiguous().to(DEV)).shape == (64 * 3, 3, 3)
    assert fe.estimate_model(f["matches"][None, :200].to(DEV)).shape == (1, 3, 3)       # non-minimal refit
    r = golden("rigid")
    model, R, t, scale = RigidTransformationSVDBasedSolver().estimate_model(r["pts"].to(DEV), flag=False)
    assert torch.allclose(model.cpu(), r["model_0"], atol=5e-4, rtol=1e-4) and scale.shape == (64,)


def test_driver_test_mode_vs_reference_loop(golden):
    """RANSAC.__call__ in test mode, two chunks of 32 with the reference's noise: the same winner."""
    from differentiable_ransac_b200.model_cl import RANSACLayer
    g = golden("driver_test")
    layer = RANSACLayer(_opt(adaptive=False))
    drv = layer.estimator
    drv.max_iterations = 64
    drv.final_refit = False
    drv.sampler.injected_noise = g["noise"].reshape(64, -1).to(DEV)
    model, mask, score, its = drv(g["matches"].to(DEV), g["logits"].to(DEV), K1, K1, None)
    assert its == 64
    rm = g["best_model"]
    assert min((model.cpu() - rm).norm(), (model.cpu() + rm).norm()) < 2e-3
    inter = (mask.cpu() & g["best_mask"]).sum().item()
    assert inter / max((mask.cpu() | g["best_mask"]).sum().item(), 1) > 0.95
    assert abs(float(score) - float(g["best_score"])) < 0.02 * float(g["best_score"])     # fp32 LAPACK noise, SURVEY H7


def test_layer_train_mode_and_match_loss(golden):
    from differentiable_ransac_b200.loss import MatchLoss
    from differentiable_ransac_b200.model_cl import RANSACLayer
    g = golden("driver_train_64")
    layer = RANSACLayer(_opt(tr=1))
    layer.estimator.max_iterations = 64
    layer.estimator.sampler.injected_noise = g["noise"].reshape(64, -1).to(DEV)
    logits = g["logits"].to(DEV).requires_grad_(True)
    Es, secs = layer(g["matches"].to(DEV), logits, K1, K1, None, None, g["E_gt"].to(DEV))
    assert Es.shape == (int(g["sel_keep"].sum()), 3, 3) and secs > 0
    m = g["matches"].to(DEV)
    loss = MatchLoss(0).forward([Es], g["E_gt"][None].numpy(), [m[:, :2]], [m[:, 2:]], None, None, None, None,
                                gt_masks=[g["gt_mask"].to(DEV)])
    loss.backward()
    d = torch.minimum((Es.detach().cpu() - g["sel_models"]).flatten(1).norm(dim=1),
                      (Es.detach().cpu() + g["sel_models"]).flatten(1).norm(dim=1))
    if (d < 1e-3).all():
        assert abs(loss.item() - g["sel_loss"].item()) < 1e-4 * g["sel_loss"].item()
        rl = g["sel_grad_logits"]
        assert (logits.grad.cpu().double() - rl).norm() / rl.norm() < 1e-4
    assert (d < 1e-3).float().mean() > 0.95


def test_layer3d_train_mode(golden):
    from differentiable_ransac_b200.model_cl import RANSACLayer3D
    g = golden("rigid_train")
    layer = RANSACLayer3D(_opt(tr=1))
    layer.estimator.max_iterations = 64
    layer.estimator.sampler.injected_noise = g["noise"].reshape(64, -1).to(DEV)
    logits = g["logits"].to(DEV).requires_grad_(True)
    Es, loss, avg_loss, _ = layer(g["points"].to(DEV), logits)
    assert Es.shape == (64, 4, 4)
    assert abs(loss.item() - g["loss"].item()) < 2e-3 * g["loss"].item()
    assert abs(float(avg_loss.detach()) - float(g["mean_residuals"].mean())) < 2e-3 * float(g["mean_residuals"].mean())
    loss.backward()
    rl = g["grad_logits"]
    assert (logits.grad.cpu() - rl).norm() / rl.norm() < 2e-2


def test_forward_batched_recovers_poses():
    from differentiable_ransac_b200 import synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    B, N = 6, 2000
    matches, E_gt, _ = synth.relative_pose_batch(B, N, seed=77, noise=2e-4)
    logits = synth.logits_regime(B, N, "L0", seed=3)
    layer = RANSACLayer(_opt())
    Kb = K1[None].expand(B, 3, 3)
    Es = layer.forward_batched(matches.to(DEV), logits.to(DEV), Kb, Kb, K=2000)
    Es = torch.stack(Es).cpu()
    err = torch.minimum((Es - E_gt).flatten(1).norm(dim=1), (Es + E_gt).flatten(1).norm(dim=1))
    easy = torch.arange(B) % 3 != 0
    assert (err[easy] < 2e-2).all()
    # the per-pair call of the reference scripts agrees on the easy pairs
    E1, _ = layer(matches[1].to(DEV), logits[1].to(DEV), K1, K1, None, None)
    # adaptive early exit (reference behaviour) stops after a few chunks of 32: a noisier minimal model
    assert min((E1.cpu() - E_gt[1]).norm(), (E1.cpu() + E_gt[1]).norm()) < 6e-2


def test_fundamental_layer_test_mode_with_refit():
    from differentiable_ransac_b200 import synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    pm, F_gt, Kc, inl = synth.pixel_pair(2000, 0.6, seed=5)
    im = torch.tensor([480.0, 640.0])
    # the layer expects image-normalised points and de-normalises them itself (model_cl.py:239-242)
    pn = pm.clone()
    pn[:, 0:2] = (pm[:, 0:2] - torch.stack((im[1] / 2, im[0] / 2))) / max(im)
    pn[:, 2:4] = (pm[:, 2:4] - torch.stack((im[1] / 2, im[0] / 2))) / max(im)
    layer = RANSACLayer(_opt(fmat=1, sampler=3, ransac_batch_size=256))
    layer.estimator.max_iterations = 1024
    logits = torch.rand(2000)
    F, _ = layer(pn.to(DEV), logits.to(DEV), Kc, Kc, im.to(DEV), im.to(DEV))
    Fu, Fg = unit(F.cpu()), unit(F_gt)
    assert min((Fu - Fg).norm(), (Fu + Fg).norm()) < 5e-2


def test_fundamental_seven_point_test_mode():
    """`-fmat 1 -sam 2`: 7-point samples (up to three models each) through the driver."""
    from differentiable_ransac_b200 import synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    pm, F_gt, Kc, inl = synth.pixel_pair(2000, 0.6, seed=6)
    im = torch.tensor([480.0, 640.0])
    pn = pm.clone()
    pn[:, 0:2] = (pm[:, 0:2] - torch.stack((im[1] / 2, im[0] / 2))) / max(im)
    pn[:, 2:4] = (pm[:, 2:4] - torch.stack((im[1] / 2, im[0] / 2))) / max(im)
    layer = RANSACLayer(_opt(fmat=1, sampler=2, ransac_batch_size=256))
    assert layer.estimator.sampler.num_samples == 7
    layer.estimator.max_iterations = 1024
    F, _ = layer(pn.to(DEV), torch.rand(2000).to(DEV), Kc, Kc, im.to(DEV), im.to(DEV))
    Fu, Fg = unit(F.cpu()), unit(F_gt)
    assert min((Fu - Fg).norm(), (Fu + Fg).norm()) < 5e-2
