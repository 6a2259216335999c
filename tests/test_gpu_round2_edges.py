"""Edge cases of the round-2 kernels: the rigid residual from second moments, the equal per-pair slices of the symmetric
epipolar loss, the float64 chain without a solution, and the raw input copies of the pipelined plugin call -- ragged,
tiny and degenerate inputs against plain fp64 torch restatements of the reference formulas
(rigid_transformation_SVD_based_solver.py:76-89, model_cl.py:13-26 + loss.py:138-144)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _rigid_reference(points, models, g_res):
    """sum_n ||q - (R p + t)||^2 per model and d(sum_k g_k res_k) / d model, in fp64 by autograd."""
    p, q = points[..., :3].double(), points[..., 3:].double()
    md = models.double().clone().requires_grad_(True)
    R, t = md[..., :3, :3], md[..., :3, 3]
    d = q[:, None] - (torch.einsum("bkij,bnj->bkni", R, p) + t[:, :, None])
    res = (d * d).sum((-1, -2))
    (res * g_res.double()).sum().backward()
    return res.detach(), md.grad


@pytest.mark.parametrize("B,K,N", [(1, 1, 1), (2, 7, 5), (3, 257, 1001), (1, 600, 4096)])
def test_rigid_residual_from_moments_any_shape(B, K, N):
    """K on both sides of the 256-model CTA, N from one point up; forward, backward and the fused entry against fp64
    autograd; the per-point kernel (asked for inlier counts) gives the same sums."""
    from differentiable_ransac_b200 import ops
    gen = torch.Generator().manual_seed(B * 1000 + K + N)
    points = torch.randn(B, N, 6, generator=gen) * 2.0 + 1.0
    models = torch.zeros(B, K, 4, 4)
    models[..., :3, :] = torch.randn(B, K, 3, 4, generator=gen)
    models[..., 3, 3] = 1.0
    g_res = torch.randn(B, K, generator=gen)
    want_res, want_g = _rigid_reference(points, models, g_res)
    pd, md, gd = points.to(DEV), models.to(DEV), g_res.to(DEV)
    res, _ = ops.rigid_residual_forward(pd, md, want_ninl=False)             # moments
    res_pp, ninl = ops.rigid_residual_forward(pd, md, want_ninl=True)        # per point (inlier counts)
    gm = ops.rigid_residual_backward(pd, md, gd).reshape(B, K, 4, 4)
    res_f, gm_f = ops.rigid_residual_forward_backward(pd, md, gd)
    scale = want_res.abs().clamp_min(1e-6)
    assert ((res.cpu().double() - want_res).abs() / scale).max() < 1e-6
    assert ((res_pp.cpu().double() - want_res).abs() / scale).max() < 1e-4     # fp32 sums over the points
    assert torch.equal(res_f, res) and torch.equal(gm_f.reshape(B, K, 4, 4), gm)
    gscale = want_g[..., :3, :].abs().amax((-1, -2)).clamp_min(1e-6)[..., None, None]
    assert ((gm.cpu().double()[..., :3, :] - want_g[..., :3, :]).abs() / gscale).max() < 1e-5
    assert (gm[..., 3, :] == 0).all()                                          # only the [R | t] block carries a gradient
    assert int(ninl.min()) >= 0


def test_rigid_residual_moments_exact_fit_does_not_cancel():
    """The case fp32 moments would get wrong: a model that fits every point, |q|^2 N ~ 1e5 against a residual sum of
    ~1e-3.  The moments are fp64, so the tiny sum survives."""
    from differentiable_ransac_b200 import ops
    gen = torch.Generator().manual_seed(3)
    N = 20000
    p = torch.randn(1, N, 3, generator=gen, dtype=torch.float64) * 2
    R = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64)).Q
    t = torch.tensor([0.3, -1.2, 2.0], dtype=torch.float64)
    q = p @ R.T + t + 1e-4 * torch.randn(1, N, 3, generator=gen, dtype=torch.float64)
    points = torch.cat((p, q), -1).float()
    model = torch.eye(4).repeat(1, 1, 1, 1)
    model[0, 0, :3, :3], model[0, 0, :3, 3] = R.float(), t.float()
    want, _ = _rigid_reference(points, model, torch.ones(1, 1))
    got, _ = ops.rigid_residual_forward(points.to(DEV), model.to(DEV), want_ninl=False)
    assert float(want) < 1e-2
    assert abs(float(got) - float(want)) < 1e-4 * float(want)


def _episym_reference(pts, npts, models, g_row):
    B, K = models.shape[:2]
    md = models.double().clone().requires_grad_(True)
    rows = []
    for b in range(B):
        P = int(npts[b])
        x = pts[b, :P].double()
        h1 = torch.cat((x[:, :2], torch.ones(P, 1, dtype=torch.float64)), -1)
        h2 = torch.cat((x[:, 2:], torch.ones(P, 1, dtype=torch.float64)), -1)
        Fx1 = torch.einsum("kij,pj->kpi", md[b], h1)
        Ftx2 = torch.einsum("kji,pj->kpi", md[b], h2)
        r = (Fx1 * h2[None]).sum(-1)
        ys = r * r * (1.0 / (Fx1[..., 0] ** 2 + Fx1[..., 1] ** 2 + 1e-15) + 1.0 / (Ftx2[..., 0] ** 2 + Ftx2[..., 1] ** 2 + 1e-15))
        rows.append(torch.clamp(ys, max=1.0).sum(-1))
    row = torch.stack(rows)
    (row * g_row.double()).sum().backward()
    return row.detach(), md.grad


def test_episym_equal_slices_on_ragged_pairs():
    """Pairs with 0, 1, 255, 256, 257 and 1500 points in one batch (slices of ~256 points, as many as each pair needs)."""
    from differentiable_ransac_b200 import ops, synth
    counts = [0, 1, 255, 256, 257, 1500]
    B, K, P = len(counts), 150, max(counts)
    matches, E_gt, _ = synth.relative_pose_batch(B, P, seed=9, noise=1e-3)
    gen = torch.Generator().manual_seed(1)
    models = torch.randn(B, K, 3, 3, generator=gen)
    models = models / models.flatten(-2).norm(dim=-1)[..., None, None]
    models[:, 0] = E_gt                                       # one model per pair that leaves the clamp
    g_row = torch.randn(B, K, generator=gen)
    npts = torch.tensor(counts, dtype=torch.int32)
    want_row, want_g = _episym_reference(matches, npts, models, g_row)
    args = (matches.to(DEV), models.to(DEV))
    row = ops.episym_forward(*args, npts.to(DEV))
    gm = ops.episym_backward(*args, g_row.to(DEV), npts.to(DEV))
    row_f, gm_f = ops.episym_forward_backward(*args, g_row.to(DEV), npts.to(DEV))
    for r_ in (row, row_f):
        assert ((r_.cpu().double() - want_row).abs() / want_row.clamp_min(1.0)).max() < 1e-4
    assert (row[0] == 0).all() and (gm[0] == 0).all()         # the pair without points
    for g_ in (gm, gm_f):
        num = (g_.cpu().double() - want_g).flatten(2).norm(dim=-1)
        den = want_g.flatten(2).norm(dim=-1)
        rel = (num / den.clamp_min(1e-12))[den > 1e-9]
        assert rel.median() < 1e-5 and rel.quantile(0.99) < 2e-3, (float(rel.median()), float(rel.max()))


def test_f64_chain_without_a_model():
    """Slots that hold no model (nsol = 0: the solver found no real root) are not scored (-1) and cannot win: the winner
    is "none" (-1) with the identity and an empty mask -- the fp32 path's convention (nister.py:400-405).  Slots beyond
    nsol are skipped even when they hold a perfectly good matrix."""
    from differentiable_ransac_b200 import ops, synth
    B, K, N = 2, 3, 200
    matches, E_gt, _ = synth.relative_pose_batch(B, N, seed=4)
    matches = matches.double()
    models = E_gt.double()[:, None, None].expand(B, K, 10, 3, 3).contiguous()      # the GT model in every slot
    thr = torch.full((B,), 1e-3, dtype=torch.float64)
    nsol = torch.zeros(B, K, dtype=torch.int32)
    scores = ops.score_msac_f64(matches.to(DEV), models.to(DEV), thr.to(DEV), nsol.to(DEV))
    assert (scores == -1).all()
    best_id, best_score, best_model, mask, ninl = ops.best_finalize_f64(matches.to(DEV), models.to(DEV), scores, thr.to(DEV))
    assert (best_id == -1).all() and (best_score == 0).all() and (mask == 0).all() and (ninl == 0).all()
    assert torch.equal(best_model.cpu(), torch.eye(3, dtype=torch.float64).expand(B, 3, 3))
    nsol[1, 2] = 4                                                                  # slots 0..3 of one hypothesis count
    scores = ops.score_msac_f64(matches.to(DEV), models.to(DEV), thr.to(DEV), nsol.to(DEV))
    assert int((scores[1] >= 0).sum()) == 4 and (scores[0] == -1).all()
    best_id, best_score, _, mask, ninl = ops.best_finalize_f64(matches.to(DEV), models.to(DEV), scores, thr.to(DEV))
    assert int(best_id[0]) == -1 and int(best_id[1]) == 20 and float(best_score[1]) > 10 and int(ninl[1]) == int(mask[1].sum()) > 10


def test_plugin_submit_takes_any_host_tensor():
    """RANSACLayer.submit copies float32 contiguous host tensors with one foreign call each; anything else -- float64,
    a strided view, unpinned memory, per-batch [3,3] intrinsics -- goes through the framework's copy: same results."""
    import types

    from differentiable_ransac_b200 import synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    B, N, K = 3, 500, 64
    m, _, _ = synth.relative_pose_batch(B, N, seed=2, noise=2e-4)
    lg = synth.logits_regime(B, N, "L0", seed=3)
    Kc = torch.tensor([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    outs = []
    for variant in range(3):
        opt = types.SimpleNamespace(device=DEV, fmat=0, sampler=2, precision=1, tr=0, threshold=0.75, ransac_batch_size=K,
                                    weighted=0, adaptive=False, final_refit=False, seed=11)
        layer = RANSACLayer(opt)
        layer.estimator.max_iterations = K
        if variant == 0:
            args = (m.pin_memory(), lg.pin_memory(), Kc.expand(B, 3, 3).contiguous(), Kc.expand(B, 3, 3).contiguous())
        elif variant == 1:
            wide = torch.zeros(B, N, 8)
            wide[..., ::2] = m                                   # a strided view of float32 memory, unpinned
            args = (wide[..., ::2], lg.double(), Kc, Kc)          # float64 weights, unbatched intrinsics
        else:
            args = (m.double(), lg, Kc.double().expand(B, 3, 3), Kc.expand(B, 3, 3))
        Es, masks, scores = layer.collect(layer.submit(*args, slots=2))
        outs.append((torch.stack([e.clone() for e in Es]), masks.clone(), scores.clone()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    assert float(outs[0][2].min()) > 10
