"""engine.TrainStep (forward + loss + backward as a fixed launch sequence, one CUDA graph) against the autograd
path it replaces (`HypothesizeE5/F8/Rigid` -> `match_loss` / `RigidResidual` -> `.backward()`, i.e. what
train.py:150-175 runs): same loss and the same d loss / d logits, step after step (the device-side Philox counter
must advance exactly like the host-side offset)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _inlier_pack(matches, inl):
    B = matches.shape[0]
    npts = inl.sum(1).int()
    P = int(npts.max())
    pts = torch.zeros(B, P, 4)
    for b in range(B):
        pts[b, : int(npts[b])] = matches[b][inl[b]]
    return pts, npts


def _autograd_step(kind, m, lg, K, seed, offset, gt=None, pts=None, npts=None):
    from differentiable_ransac_b200 import engine
    m = m.clone().requires_grad_(True)
    lg = lg.clone().requires_grad_(True)
    if kind == "e5":
        models, valid = engine.HypothesizeE5.apply(m, lg, gt, K, 1.0, None, seed, offset, True)
        loss = engine.match_loss(models, valid, pts, npts).mean()
    elif kind == "f8":
        models, valid = engine.HypothesizeF8.apply(m, lg, K, 1.0, None, seed, offset)
        loss = engine.match_loss(models, valid, pts, npts).mean()
    else:
        models, valid = engine.HypothesizeRigid.apply(m, lg, K, True, 1.0, None, seed, offset)
        res = engine.RigidResidual.apply(m, torch.where(valid[..., None, None], models, torch.zeros_like(models)))
        v = valid.float()
        loss = ((res * v).sum(1) / (v.sum(1).clamp_min(1.0) * m.shape[1])).mean()
    loss.backward()
    return loss.detach(), lg.grad, m.grad


@pytest.mark.parametrize("kind", ["e5", "f8", "rigid"])
@pytest.mark.parametrize("graph", [False, True])
def test_train_step_equals_the_autograd_path(kind, graph):
    from differentiable_ransac_b200 import engine, synth
    B, N, K, seed = 3, 600, 96, 11
    gt = pts = npts = None
    if kind == "e5":
        m, gt, inl = synth.relative_pose_batch(B, N, seed=31, noise=2e-4)
        pts, npts = _inlier_pack(m, inl)
        gt = gt.to(DEV)
    elif kind == "f8":
        pairs = [synth.pixel_pair(N, 0.5, seed=40 + b) for b in range(B)]
        m = torch.stack([p[0] / 640.0 for p in pairs])
        pts, npts = _inlier_pack(m, torch.stack([p[3] for p in pairs]))
    else:
        m = torch.stack([synth.rigid_pair(N, 0.7, seed=50 + b)[0] for b in range(B)])
    lg = synth.logits_regime(B, N, "L0" if kind == "e5" else "L1", seed=3)
    m, lg = m.to(DEV), lg.to(DEV)
    if pts is not None:
        pts, npts = pts.to(DEV), npts.to(DEV)
    step = engine.TrainStep(kind, B, N, K, DEV, P=None if pts is None else pts.shape[1], seed=seed, graph=graph,
                            want_grad_matches=True)
    for it in range(3):                      # the device-side counter must track the host-side offset
        loss, gl = step.run(m, lg, gt, pts, npts)
        gm = step.grad_matches
        torch.cuda.synchronize()
        want_loss, want_gl, want_gm = _autograd_step(kind, m, lg, K, seed, it, gt, pts, npts)
        assert torch.isfinite(gl).all() and torch.isfinite(loss).all()
        assert abs(float(loss) - float(want_loss)) <= 1e-6 * max(1.0, abs(float(want_loss)))
        scale = float(want_gl.abs().max())
        assert scale > 0
        assert float((gl - want_gl).abs().max()) <= 1e-5 * scale
        assert float((gm - want_gm).abs().max()) <= 1e-5 * max(float(want_gm.abs().max()), 1e-30)
    # successive steps drew different hypotheses
    a, _ = step.run(m, lg, gt, pts, npts)
    a = float(a)
    b, _ = step.run(m, lg, gt, pts, npts)
    assert a != float(b)


def test_fused_entries_equal_their_parts():
    """drb_solve_e5_select = drb_solve_e5 + drb_select_closest; drb_episym_forward_backward and
    drb_rigid_residual_forward_backward = their forward and backward launches (one pass over the points instead of
    two); drb_solve_e5_backward_chosen = drb_solve_e5_backward."""
    from differentiable_ransac_b200 import ops, synth
    B, N, K = 3, 700, 150
    m, gt, inl = synth.relative_pose_batch(B, N, seed=61, noise=2e-4)
    pts, npts = _inlier_pack(m, inl)
    lg = synth.logits_regime(B, N, "L0", seed=5)
    m, gt, lg, pts, npts = m.to(DEV), gt.to(DEV), lg.to(DEV), pts.to(DEV), npts.to(DEV)
    idx = ops.sample_sets(lg, K, 5, seed=3)
    models, nsol = ops.solve_e5(m, idx)
    for si in (True, False):
        sel0, ch0 = ops.select_closest(models, nsol, gt, si)
        sel1, ch1, nsol1, dense = ops.solve_e5_select(m, idx, gt, si, want_models=True)
        assert torch.equal(sel0, sel1) and torch.equal(ch0, ch1) and torch.equal(nsol, nsol1) and torch.equal(dense, models)
    sel, chosen, _, none = ops.solve_e5_select(m, idx, gt, True)
    assert none is None and torch.equal(chosen, ch0 if False else ops.select_closest(models, nsol, gt, True)[1])
    valid = sel >= 0
    g_row = torch.rand(B, K, device=DEV) * valid
    row0 = ops.episym_forward(pts, chosen, npts, valid)
    g0 = ops.episym_backward(pts, chosen, g_row, npts, valid)
    row1, g1 = ops.episym_forward_backward(pts, chosen, g_row, npts, valid)
    assert torch.allclose(row0, row1, rtol=1e-6, atol=1e-6) and torch.allclose(g0, g1, rtol=1e-6, atol=1e-9)
    gm = torch.randn(B, K, 9, device=DEV)
    assert torch.equal(ops.solve_e5_backward(m, idx, models, sel, gm), ops.solve_e5_backward_chosen(m, idx, chosen, sel, gm))
    rp = torch.stack([synth.rigid_pair(1500, 0.7, seed=70 + b)[0] for b in range(2)]).to(DEV)
    l3 = synth.logits_regime(2, 1500, "L1", seed=2).to(DEV)
    rm, rv = ops.solve_rigid3(rp, ops.sample_sets(l3, 40, 3, seed=1), True)
    rm = torch.where(rv.bool()[..., None, None], rm, torch.zeros_like(rm))
    g_res = torch.rand(2, 40, device=DEV)
    r0, _ = ops.rigid_residual_forward(rp, rm, want_ninl=False)
    gg0 = ops.rigid_residual_backward(rp, rm, g_res)
    r1, gg1 = ops.rigid_residual_forward_backward(rp, rm, g_res)
    assert torch.allclose(r0, r1, rtol=1e-5) and torch.allclose(gg0, gg1, rtol=1e-6, atol=1e-9)


def test_episym_forward_zeroes_the_rows_of_invalid_models():
    """Regression: with a validity mask and one split the rows of invalid models were left as allocated
    (torch.empty), so `row * valid` in engine.match_loss could be 0 * NaN."""
    from differentiable_ransac_b200 import ops
    B, K, P = 2, 96, 300
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(B, P, 4, generator=g) - 0.5).to(DEV)
    models = torch.randn(B, K, 3, 3, generator=g).to(DEV)
    valid = (torch.rand(B, K, generator=g) < 0.5).to(DEV)
    for _ in range(4):
        poison = torch.full((B, K), float("nan"), device=DEV)    # the block the next torch.empty(B, K) gets
        del poison
        row = ops.episym_forward(pts, models, None, valid)
        assert torch.isfinite(row).all()
        assert (row[~valid] == 0).all() and (row[valid] > 0).all()
