"""The CPU oracle against the reference's own outputs (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference).  CPU only."""
import pytest
import torch

from oracle import driver, fundamental, nister, rigid, sampler, scoring, stewenius
from helpers import match_up_to_sign, trace_constraint_residual, unit


def test_sampler_and_gather(golden):
    for regime in ("L0", "L1"):
        g = golden(f"sampler_{regime}")
        ret, y_soft, idx = sampler.sample(g["logits"], g["noise"], 5)
        assert torch.equal(idx, g["idx"])
        assert torch.equal(ret, g["ret"])
        assert torch.equal(y_soft, g["y_soft"])
        assert torch.equal(sampler.gather_minimal(g["matches"], ret), g["minimal"])
        # the ascending-index gather equals plain indexing (straight-through value is exactly 1)
        assert torch.equal(g["matches"][idx], g["minimal"])


def test_nister_fp64_models_match_reference(golden):
    g = golden("nister")
    E = nister.five_point(g["pts"].double())
    assert E.shape == g["E64"].shape
    real = trace_constraint_residual(g["E64"]) < 1e-8
    # same LAPACK calls on the same inputs: slot-for-slot agreement on the genuine models
    assert (E[real] - g["E64"][real]).abs().max() < 1e-6


def test_nister_fp32_within_reference_noise_floor(golden):
    g = golden("nister")
    E32 = nister.five_point(g["pts"])
    ref64 = g["E64"].view(-1, 10, 3, 3)
    real = (trace_constraint_residual(g["E64"]) < 1e-8).view(-1, 10)
    d_or = match_up_to_sign(E32.reshape(-1, 10, 3, 3), ref64)[real]
    d_ref = match_up_to_sign(g["E32"].reshape(-1, 10, 3, 3), ref64)[real]
    # the oracle in fp32 is as close to the fp64 reference as the fp32 reference is
    assert (d_or < 1e-3).float().mean() >= (d_ref < 1e-3).float().mean() - 0.03
    assert d_or.median() < 1e-4


def test_stewenius_fp64_and_the_two_classes_define_the_same_models(golden):
    """The reference's Stewenius class run in fp64 (stewenius_64.npz): the port reproduces it, and its genuine models
    (trace-constraint residual < 1e-8) are, up to scale and sign, exactly the genuine models of the reference's
    Nister class in fp64 -- same count, every model matched to 1e-7.  This is what lets ONE device kernel stand
    behind both estimator classes."""
    from helpers import match_up_to_sign, trace_constraint_residual, unit
    g, g64 = golden("stewenius"), golden("stewenius_64")
    pts = g["pts"].double()
    K = pts.shape[0]
    ref = unit(g64["E64"]).view(K, 10, 3, 3)
    port = unit(stewenius.five_point(pts)).view(K, 10, 3, 3)
    real = (trace_constraint_residual(unit(g64["E64"])) < 1e-8).view(K, 10)
    assert float(match_up_to_sign(port, ref)[real].max()) < 1e-10
    En = nister.five_point(pts).view(K, 10, 3, 3)
    gn = (trace_constraint_residual(En.reshape(-1, 3, 3)) < 1e-8).view(K, 10)
    assert int(real.sum()) == int(gn.sum()) and torch.equal(real.sum(1), gn.sum(1))
    assert float(match_up_to_sign(unit(En), ref)[real].max()) < 1e-7
    assert float(match_up_to_sign(ref, unit(En))[gn].max()) < 1e-7


def test_stewenius(golden):
    g = golden("stewenius")
    E = stewenius.five_point(g["pts"])
    assert torch.allclose(E, g["E32"], atol=1e-5)


def test_eight_point(golden):
    g = golden("f8")
    assert torch.allclose(fundamental.eight_point(g["pts"].double()), g["F64"], atol=1e-9)
    d = (unit(fundamental.eight_point(g["pts"])) - unit(g["F32"])).flatten(1).norm(dim=1)
    assert d.max() < 1e-4


def test_seven_point_algebra(golden):
    g = golden("f8")
    pts = g["pts"][:, :7].double()
    F, valid = fundamental.seven_point(pts)
    assert valid.any(dim=1).all()
    h1 = torch.cat((pts[..., :2], torch.ones_like(pts[..., :1])), -1)
    h2 = torch.cat((pts[..., 2:], torch.ones_like(pts[..., :1])), -1)
    for s in range(3):
        v = valid[:, s]
        r = torch.einsum("kni,kij,knj->kn", h2[v], F[v, s], h1[v])
        assert r.abs().max() < 1e-6
        assert torch.linalg.det(F[v, s]).abs().max() < 1e-10


def test_rigid(golden):
    g = golden("rigid")
    for flag in (True, False):
        m, R, t, s = rigid.estimate(g["pts"], flag=flag)
        assert torch.allclose(m, g[f"model_{int(flag)}"], atol=1e-6)
        r, mr, mask = scoring.rigid_squared_residual(g["points"][:, :3], g["points"][:, 3:],
                                                     m[:, :3, :].transpose(-1, -2))
        assert torch.allclose(r, g[f"res_{int(flag)}"], rtol=1e-6)
        assert torch.equal(mask.sum(-1), g[f"ninl_{int(flag)}"])


def test_msac_and_episym(golden):
    g = golden("msac")
    sc, masks = scoring.msac_score(g["matches"], g["models"], float(g["threshold"]))
    # matmul blocking depends on the host thread count, so not bit-equal
    assert torch.allclose(sc, g["scores"], rtol=1e-4)
    assert int(torch.argmax(sc)) == int(g["best"])
    assert torch.equal(masks[int(g["best"])], g["best_mask"])
    e = golden("episym")
    loss = scoring.match_loss(e["models"], e["matches"][:, :2], e["matches"][:, 2:], e["gt_mask"])
    assert torch.allclose(loss, e["loss"], rtol=1e-6)


def test_driver_test_loop(golden):
    g = golden("driver_test")
    thr = driver.normalized_threshold(0.75, g["K1"], g["K1"], fmat=False)
    assert abs(thr - float(g["threshold"])) < 1e-12
    out = driver.test_loop(g["matches"], g["logits"], list(g["noise"]), thr)
    assert out["best_chunk"] == int(g["best_chunk"])
    assert out["best_idx"] // 10 == int(g["best_hyp"])
    for c in range(2):
        assert torch.equal(out["samples"][c], g["idx"][c])


def test_driver_train_loop_fp64(golden):
    g = golden("driver_train_64")
    Es = driver.train_loop(g["matches"].double(), g["logits"].double(), list(g["noise"].double()),
                           g["E_gt"].double())
    assert torch.allclose(Es, g["models"], atol=1e-6)


def test_rigid_train_loop(golden):
    g = golden("rigid_train")
    models, res, mres = driver.rigid_train_loop(g["points"], g["logits"], list(g["noise"]))
    assert torch.allclose(torch.cat(models), g["models"], atol=1e-6)
    assert torch.allclose(torch.cat(res), g["residuals"], rtol=1e-5)


def test_nonminimal_fits(golden):
    """the fits behind the final refit and LO (nister.py:51-65, fundamental...:169-175)."""
    g = golden("refit_e5")
    m, mask = g["matches"], g["mask"].bool()
    assert torch.allclose(nister.five_point(m[None].double()), g["E_all64"], atol=1e-9)
    assert torch.allclose(nister.five_point(m[mask][None].double()), g["E_inl64"], atol=1e-9)
    f = golden("refit_f8")
    m, mask, w = f["matches"], f["mask"].bool(), f["weights"]
    assert torch.allclose(fundamental.eight_point(m[mask][None].double()), f["F_inl64"], atol=1e-12)
    assert torch.allclose(fundamental.eight_point(m[mask][None].double(), w[mask][None].double()), f["F_w64"],
                          atol=1e-12)


@pytest.mark.parametrize("name,fmat,s", [("driver_full_e5_lo0", False, 5), ("driver_full_e5_lo2", False, 5),
                                         ("driver_full_e5_lo0_64", False, 5), ("driver_full_e5_lo2_64", False, 5),
                                         ("driver_full_f8_lo0", True, 8), ("driver_full_f8_lo2", True, 8)])
def test_full_test_driver(golden, name, fmat, s):
    """`RANSAC.__call__` in test mode run by the reference itself: adaptive exit, LO (lo=2, 8 iterations), final
    refit."""
    from differentiable_ransac_b200 import synth
    g = golden(name)
    dt = torch.float64 if name.endswith("_64") else torch.float32
    m = g["matches"].to(dt)
    Kc = (g["K"] if fmat else g["K1"]).to(dt)
    noises = [synth.gumbel_noise((32, m.shape[0]), seed=int(sd)).to(dt) for sd in g["noise_seeds"]]
    model, mask, score, its = driver.full_test_driver(m, g["logits"].to(dt), noises, Kc, Kc, 0.75, fmat=fmat,
                                                      sample_size=s, lo=int(name.split("lo")[1][0]), lo_iters=8)
    assert its == int(g["iterations"])
    assert abs(float(score) - float(g["best_score"])) < 1e-3 * float(g["best_score"])
    assert min((model - g["best_model"]).abs().max(), (model + g["best_model"]).abs().max()) < 1e-4
    assert (mask != g["best_mask"].bool()).sum() <= 1


def test_pose_recovery_restatement(golden):
    """`cv_utils.recoverPose` / `eval_essential_matrix` and OpenCV's recoverPose mask as
    MatchLoss uses it (rank 3)."""
    from oracle import pose_eval
    g = golden("pose")
    for i in range(int(g["n_cases"])):
        m = g[f"matches_{i}"].numpy()
        for j in range(3):
            R, t, mask, _ = pose_eval.recover_pose_ref(g[f"E_{i}"][j].numpy(), m[:, :2], m[:, 2:])
            assert abs(R - g[f"R_{i}"][j].numpy()).max() < 1e-9 and abs(t - g[f"t_{i}"][j].numpy()).max() < 1e-9
            er, et = pose_eval.pose_error_ref(R, t, g[f"R_gt_{i}"].numpy(), g[f"t_gt_{i}"].numpy())
            assert abs(er - float(g[f"err_{i}"][j, 0])) < 1e-5 and abs(et - float(g[f"err_{i}"][j, 1])) < 1e-5
            if j == 0:
                assert (mask == g[f"cv_mask_{i}"].numpy().astype(bool)).all()
