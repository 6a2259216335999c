"""Edge cases of the CUDA path: degenerate / non-finite minimal samples, tiny and large shapes,
ragged counts, error codes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _finite_or_flagged(models, nvalid, slots):
    """Every slot the kernel calls valid must be finite; the rest must be the identity."""
    models = models.reshape(-1, slots, 3, 3).cpu()
    nvalid = nvalid.reshape(-1).cpu()
    live = torch.arange(slots)[None] < nvalid[:, None]
    assert torch.isfinite(models[live]).all()
    assert torch.equal(models[~live], torch.eye(3).expand(int((~live).sum()), 3, 3))


def test_degenerate_five_point_samples():
    from differentiable_ransac_b200 import ops
    g = torch.Generator().manual_seed(0)
    base = torch.rand(6, 5, 4, generator=g) - 0.5
    same = base[0:1, 0:1].expand(1, 5, 4)                                   # five identical correspondences
    line = torch.stack([torch.tensor([t, 2 * t, t, 2 * t]) for t in torch.linspace(-0.4, 0.4, 5)])[None]
    zero = torch.zeros(1, 5, 4)
    nan = base[1:2].clone()
    nan[0, 2, 1] = float("nan")
    inf = base[2:3].clone()
    inf[0, 0, 0] = float("inf")
    huge = base[3:4] * 1e18
    pts = torch.cat((same, line, zero, nan, inf, huge, base)).to(DEV)
    models, nsol = ops.solve_e5(pts)
    _finite_or_flagged(models, nsol, 10)
    assert int(nsol[0, 3]) == 0 and int(nsol[0, 4]) == 0                      # NaN / Inf in -> no models out
    assert (nsol[0, 6:] > 0).any()                                            # the regular samples still solve
    # scoring such a list never lets a degenerate model win
    m = torch.rand(1, 500, 4, device=DEV) - 0.5
    sc, best = ops.score_msac(m, models.reshape(1, -1, 9), torch.tensor([1e-3], device=DEV))
    assert torch.isfinite(sc).all()


def test_degenerate_eight_point_and_rigid_samples():
    from differentiable_ransac_b200 import ops
    same8 = torch.full((1, 8, 4), 0.25)
    nan8 = torch.rand(1, 8, 4)
    nan8[0, 3, 2] = float("nan")
    F, valid = ops.solve_f8(torch.cat((same8, nan8, torch.rand(2, 8, 4))).to(DEV))
    valid = valid[0].cpu().bool()
    assert not valid[0] and not valid[1] and valid[2:].all()
    assert torch.isfinite(F[0][valid]).all()
    same3 = torch.ones(1, 3, 6)
    col = torch.stack([torch.tensor([t, t, t, 2 * t, 2 * t, 2 * t]) for t in (0.0, 1.0, 2.0)])[None]   # collinear
    for flag in (True, False):
        M, ok = ops.solve_rigid3(torch.cat((same3, col, torch.rand(3, 3, 6))).to(DEV), flag=flag)
        ok = ok[0].cpu().bool()
        assert not ok[0] and ok[2:].all()
        if not flag:
            assert not ok[1]                                                    # rank-1 covariance: no rotation
        assert torch.isfinite(M[0][ok]).all()


def test_tiny_shapes_and_error_codes():
    from differentiable_ransac_b200 import _lib, engine, ops, synth
    m, E, _ = synth.relative_pose_batch(1, 5, seed=1)                           # N == s
    out = engine.ransac_e5_test(m.to(DEV), torch.zeros(1, 5, device=DEV), 1, torch.tensor([1e-3], device=DEV),
                                want_scores=True)
    assert out["idx"].cpu().tolist() == [[[0, 1, 2, 3, 4]]]
    assert int(out["ninl"][0]) >= 0 and out["mask"].shape == (1, 5)
    with pytest.raises(_lib.DrbError, match="bad shape"):
        ops.sample_sets(torch.zeros(1, 4, device=DEV), 3, 5)                    # s > N
    with pytest.raises(_lib.DrbError, match="unsupported"):
        ops.sample(torch.zeros(1, 40, device=DEV), 3, 6, want_lse=True)         # sample size 6
    with pytest.raises(ValueError):
        ops.solve_e5(torch.zeros(4, 6, 4, device=DEV))                          # wrong minimal-sample shape
    # a pair whose every hypothesis fails: identity model, id -1, empty mask
    bad = torch.full((1, 64, 4), float("nan"), device=DEV)
    o = engine.ransac_e5_test(bad, torch.zeros(1, 64, device=DEV), 8, torch.tensor([1e-3], device=DEV))
    assert int(o["best_id"][0]) == -1 and torch.equal(o["best_model"][0].cpu(), torch.eye(3)) and int(o["ninl"][0]) == 0


def test_reference_default_iteration_count():
    """Test mode of the reference runs up to 5000 hypotheses per pair (model_cl.py:216-219)."""
    from differentiable_ransac_b200 import engine, synth
    B, K, N = 4, 5000, 2000
    m, E, _ = synth.relative_pose_batch(B, N, seed=9, noise=2e-4)
    lg = synth.logits_regime(B, N, "L0", seed=1)
    thr = torch.full((B,), 0.75 / 800)
    o = engine.ransac_e5_test(m.to(DEV), lg.to(DEV), K, thr.to(DEV))
    bm = o["best_model"].cpu()
    err = torch.minimum((bm - E).flatten(1).norm(dim=1), (bm + E).flatten(1).norm(dim=1))
    # pairs 1, 2 have 40 / 60 % inliers (dozens of all-inlier samples among 5000); pairs 0, 3 have 20 %
    # (0.2^5 * 5000 = 1.6 expected all-inlier samples: not guaranteed)
    assert (err[1:3] < 2e-2).all()
    assert (o["ninl"].cpu()[1:3] > 0.3 * N).all()


def test_unaligned_views_are_handled():
    """ops clones tensors whose storage offset breaks the 16-byte alignment the bulk copies need."""
    from differentiable_ransac_b200 import ops
    from oracle import scoring
    raw = torch.rand(1 + 300 * 4, device=DEV) - 0.5
    m = raw[1:].view(1, 300, 4)                                               # 4-byte aligned view
    models = torch.nn.functional.normalize(torch.randn(1, 40, 9, device=DEV), dim=-1)
    sc, _ = ops.score_msac(m, models, torch.tensor([0.05], device=DEV))
    ref, _ = scoring.msac_score(m[0].cpu(), models[0].cpu().view(-1, 3, 3), 0.05)
    assert torch.allclose(sc[0].cpu(), ref, rtol=1e-4, atol=1e-4)
