"""CPU checks of the DEVICE math headers (csrc/*_math.cuh) compiled for the host by
tests/hostcheck -- the same templates the CUDA kernels instantiate -- against the oracle and
the reference-generated golden fixtures.  No GPU needed; test infrastructure only."""
import ctypes
import os

import numpy as np
import pytest
import torch

import hostcheck
from helpers import match_up_to_sign, trace_constraint_residual, unit
from oracle import fundamental, nister, rigid


@pytest.fixture(scope="module")
def lib():
    return hostcheck.load()


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_e5(lib, pts, dt, polish=2):
    p = np.ascontiguousarray(pts.numpy().astype(dt))
    K = p.shape[0]
    out = np.zeros((K, 10, 9), dtype=dt)
    ns = np.zeros(K, dtype=np.int32)
    fn = lib.hc_e5_solve_f32 if dt == np.float32 else lib.hc_e5_solve_f64
    fn(vp(p), K, vp(out), vp(ns), polish)
    return torch.from_numpy(out).view(K, 10, 3, 3), torch.from_numpy(ns)


def test_e5_fp64_finds_every_reference_model(lib, golden):
    g = golden("nister")
    E, ns = run_e5(lib, g["pts"], np.float64)
    ref = g["E64"].view(-1, 10, 3, 3)
    real = (trace_constraint_residual(g["E64"]) < 1e-8).view(-1, 10)
    d = match_up_to_sign(E, ref)[real]
    assert d.max() < 1e-6
    assert (ns - real.sum(1)).abs().max() <= 1


def test_e5_fp32_beats_reference_fp32_noise_floor(lib, golden):
    g = golden("nister")
    E, ns = run_e5(lib, g["pts"], np.float32)
    ref = g["E64"].view(-1, 10, 3, 3)
    real = (trace_constraint_residual(g["E64"]) < 1e-8).view(-1, 10)
    d = match_up_to_sign(E, ref)[real]
    d_ref = match_up_to_sign(g["E32"].view(-1, 10, 3, 3), ref)[real]
    assert (d < 1e-3).float().mean() > 0.95
    assert (d < 1e-3).float().mean() >= (d_ref < 1e-3).float().mean()
    assert d.median() < 1e-5


def test_sturm_roots_against_numpy(lib):
    rng = np.random.default_rng(0)
    for trial in range(200):
        n_real = 2 * rng.integers(0, 6)
        roots = list(rng.uniform(-3, 3, n_real))
        while len(roots) < 10:
            re, im = rng.uniform(-2, 2), rng.uniform(0.2, 2)
            roots += [complex(re, im), complex(re, -im)]
        c = np.real(np.poly(roots))[::-1].copy()          # ascending
        c *= rng.uniform(0.1, 10)
        out = np.zeros(10)
        n = lib.hc_roots_f64(vp(np.ascontiguousarray(c)), vp(out))
        want = np.sort(np.real([r for r in roots if abs(np.imag(r)) < 1e-12]))
        assert n == len(want)
        assert np.allclose(np.sort(out[:n]), want, atol=1e-7)


def test_e5_backward_matches_autograd_of_the_oracle(lib, golden):
    g = golden("nister")
    gen = torch.Generator().manual_seed(0)
    errs = []
    for k in range(24):
        pts = g["pts"][k:k + 1].double().clone().requires_grad_(True)
        E = nister.five_point(pts)
        real = trace_constraint_residual(E.detach()) < 1e-9
        for s in range(10):
            if not real[s]:
                continue
            gE = torch.randn(3, 3, dtype=torch.float64, generator=gen)
            (gr,) = torch.autograd.grad((E[s] * gE).sum(), pts, retain_graph=True)
            out = np.zeros((5, 4))
            ok = lib.hc_e5_backward_f64(vp(np.ascontiguousarray(pts.detach().numpy()[0])),
                                        vp(np.ascontiguousarray(E[s].detach().numpy())),
                                        vp(np.ascontiguousarray(gE.numpy())), vp(out))
            assert ok == 1
            errs.append(np.abs(out - gr[0].numpy()).max() / np.abs(gr[0].numpy()).max())
    errs = np.array(errs)
    assert np.median(errs) < 1e-9 and (errs < 1e-4).mean() > 0.97


def test_f8_forward_backward(lib, golden):
    g = golden("f8")
    pts = np.ascontiguousarray(g["pts"].numpy().astype(np.float32))
    K = pts.shape[0]
    F = np.zeros((K, 9), np.float32)
    ok = np.zeros(K, np.int32)
    lib.hc_f8_solve_f32(vp(pts), K, vp(F), vp(ok))
    F = torch.from_numpy(F).view(K, 3, 3).double()
    R = g["F64"]
    d = torch.minimum((F - R).flatten(1).norm(dim=1), (F + R).flatten(1).norm(dim=1)) / R.flatten(1).norm(dim=1)
    assert ok.all() and d.max() < 1e-4
    p64 = g["pts"][:16].double().clone().requires_grad_(True)
    Fo = fundamental.eight_point(p64)
    gF = torch.randn(16, 3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    pn = np.ascontiguousarray(p64.detach().numpy())
    Fh = np.zeros((16, 9))
    ok = np.zeros(16, np.int32)
    lib.hc_f8_solve_f64(vp(pn), 16, vp(Fh), vp(ok))
    sgn = torch.sign((torch.from_numpy(Fh).view(16, 3, 3) * Fo.detach()).flatten(1).sum(1))
    (gr,) = torch.autograd.grad((Fo * gF).sum(), p64)
    gours = np.zeros((16, 8, 4))
    lib.hc_f8_backward_f64(vp(pn), vp(np.ascontiguousarray((gF * sgn[:, None, None]).numpy())), 16, vp(gours), vp(ok))
    rel = (torch.from_numpy(gours) - gr).flatten(1).abs().max(1).values / gr.flatten(1).abs().max(1).values
    assert ok.all() and rel.max() < 1e-7


def test_f7_against_oracle(lib, golden):
    g = golden("f8")
    pts7 = np.ascontiguousarray(g["pts"][:, :7].numpy().astype(np.float64))
    K = pts7.shape[0]
    F7 = np.zeros((K, 3, 9))
    n7 = np.zeros(K, np.int32)
    lib.hc_f7_solve_f64(vp(pts7), K, vp(F7), vp(n7))
    Fo, valid = fundamental.seven_point(torch.from_numpy(pts7))
    assert (torch.from_numpy(n7) == valid.sum(1)).all()
    F7 = torch.from_numpy(F7).view(K, 3, 3, 3)
    for k in range(K):
        for s in range(n7[k]):
            cand = Fo[k][valid[k]]
            d = torch.minimum((cand - F7[k, s]).flatten(1).norm(dim=1), (cand + F7[k, s]).flatten(1).norm(dim=1))
            assert d.min() < 1e-7


@pytest.mark.parametrize("flag", [1, 0])
def test_rigid_forward_backward(lib, golden, flag):
    g = golden("rigid")
    pts = np.ascontiguousarray(g["pts"].numpy().astype(np.float32))
    K = pts.shape[0]
    M = np.zeros((K, 16), np.float32)
    ok = np.zeros(K, np.int32)
    lib.hc_rigid3_solve_f32(vp(pts), K, flag, vp(M), vp(ok))
    assert ok.all()
    assert torch.allclose(torch.from_numpy(M).view(K, 4, 4), g[f"model_{flag}"], atol=5e-4, rtol=1e-4)
    p64 = g["pts"][:16].double().clone().requires_grad_(True)
    m, _, _, _ = rigid.estimate(p64, flag=bool(flag))
    gM = torch.randn(16, 4, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    gM[:, 3, :] = 0
    (gr,) = torch.autograd.grad((m * gM).sum(), p64)
    gours = np.zeros((16, 3, 6))
    ok = np.zeros(16, np.int32)
    lib.hc_rigid3_backward_f64(vp(np.ascontiguousarray(p64.detach().numpy())), vp(np.ascontiguousarray(gM.numpy())),
                               16, flag, vp(gours), vp(ok))
    rel = (torch.from_numpy(gours) - gr).flatten(1).abs().max(1).values / gr.flatten(1).abs().max(1).values
    assert ok.all() and rel.max() < 1e-8


def test_rigid_residual_from_moments(lib, golden):
    """The moments form of the rigid residual (rigid_math.cuh, used by score.cu's rigid_residual_moments_kernel):
    against the golden of squared_residual (rigid_transformation_SVD_based_solver.py:76-89) and against fp64 autograd
    of the point-by-point sum, including a model that fits every point (where fp32 moments would cancel)."""
    g = golden("rigid")
    for flag in (0, 1):
        pts = np.ascontiguousarray(g["points"].numpy().astype(np.float32))
        models = np.ascontiguousarray(g[f"model_{flag}"].reshape(-1, 16).numpy().astype(np.float32))
        K, N = models.shape[0], pts.shape[0]
        res, grad = np.zeros(K), np.zeros((K, 12))
        lib.hc_rigid_residual_moments(vp(pts), N, vp(models), K, vp(res), vp(grad))
        assert torch.allclose(torch.from_numpy(res).float(), g[f"res_{flag}"], rtol=1e-4)
        md = torch.from_numpy(models).double().view(K, 4, 4).clone().requires_grad_(True)
        p, q = torch.from_numpy(pts[:, :3]).double(), torch.from_numpy(pts[:, 3:]).double()
        d = q[None] - (torch.einsum("kij,nj->kni", md[:, :3, :3], p) + md[:, None, :3, 3])
        want = (d * d).sum((-1, -2))
        want.sum().backward()
        assert (torch.from_numpy(res) - want.detach()).abs().max() <= 1e-9 * want.detach().abs().max()
        wg = md.grad[:, :3, :].reshape(K, 12)
        assert (torch.from_numpy(grad) - wg).abs().max() <= 1e-9 * wg.abs().max()
    # exact fit: sum |q|^2 ~ 1e5, residual sum ~ 6e-4
    gen = torch.Generator().manual_seed(3)
    N = 20000
    p = torch.randn(N, 3, generator=gen, dtype=torch.float64) * 2
    R = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64)).Q
    t = torch.tensor([0.3, -1.2, 2.0], dtype=torch.float64)
    q = p @ R.T + t + 1e-4 * torch.randn(N, 3, generator=gen, dtype=torch.float64)
    pts = np.ascontiguousarray(torch.cat((p, q), -1).float().numpy())
    model = np.eye(4, dtype=np.float32)
    model[:3, :3], model[:3, 3] = R.float().numpy(), t.float().numpy()
    res, grad = np.zeros(1), np.zeros((1, 12))
    lib.hc_rigid_residual_moments(vp(pts), N, vp(np.ascontiguousarray(model.reshape(1, 16))), 1, vp(res), vp(grad))
    pp, qq = torch.from_numpy(pts[:, :3]).double(), torch.from_numpy(pts[:, 3:]).double()
    want = float(((qq - (pp @ torch.from_numpy(model[:3, :3]).double().T + torch.from_numpy(model[:3, 3]).double())) ** 2).sum())
    # nine orders of magnitude cancel in fp64: 1e-4 relative is what is left (fp32 moments would leave nothing)
    assert want < 1e-2 and abs(res[0] - want) < 1e-4 * want


def run_refit(lib, fmat, matches, mask=None, weights=None):
    m = np.ascontiguousarray(matches.numpy(), dtype=np.float32)
    out = np.zeros((10, 9), np.float32)
    mk = None if mask is None else np.ascontiguousarray(mask.numpy(), dtype=np.uint8)
    w = None if weights is None else np.ascontiguousarray(weights.numpy(), dtype=np.float32)
    n = lib.hc_refit(int(fmat), vp(m), None if mk is None else vp(mk), None if w is None else vp(w), m.shape[0],
                     vp(out))
    return torch.from_numpy(out[:n]).view(-1, 3, 3)


def test_refit_e5_finds_every_genuine_reference_model(lib, golden):
    """The non-minimal five-point of the final refit / LO (nister.py:51-65 on n > 5 rows), fp64 reference."""
    g = golden("refit_e5")
    for ref, mask in ((g["E_all64"], None), (g["E_inl64"], g["mask"])):
        ours = run_refit(lib, False, g["matches"], mask)
        genuine = trace_constraint_residual(ref) < 1e-8           # the other slots are complex-root leftovers (D3)
        assert int(genuine.sum()) == ours.shape[0] > 0
        d = match_up_to_sign(ours[None], unit(ref[genuine])[None])[0]
        assert d.max() < 1e-6


def test_refit_f8_matches_reference(lib, golden):
    g = golden("refit_f8")
    for ref, w in ((g["F_inl64"], None), (g["F_w64"], g["weights"])):
        F = run_refit(lib, True, g["matches"], g["mask"], w)[0].double()
        assert min((F - ref[0]).abs().max(), (F + ref[0]).abs().max()) < 1e-6   # same scale, sign arbitrary
    # fewer than eight selected correspondences: no model
    few = torch.zeros(2000, dtype=torch.bool)
    few[:7] = True
    assert run_refit(lib, True, g["matches"], few).shape[0] == 0


def test_pose_recovery_math_against_reference(lib, golden):
    """csrc/pose_math.cuh on the host: R, t, errors of `cv_utils.recoverPose` / `eval_essential_matrix` and the
    cheirality mask of cv2.recoverPose, from the reference-generated fixture."""
    g = golden("pose")
    for i in range(int(g["n_cases"])):
        m = np.ascontiguousarray(g[f"matches_{i}"].numpy(), dtype=np.float32)
        N = m.shape[0]
        Rg = np.ascontiguousarray(g[f"R_gt_{i}"].numpy(), dtype=np.float64)
        tg = np.ascontiguousarray(g[f"t_gt_{i}"].numpy(), dtype=np.float64)
        for j in range(3):
            E = np.ascontiguousarray(g[f"E_{i}"][j].numpy(), dtype=np.float64)
            R, t, err = np.zeros(9), np.zeros(3), np.zeros(2)
            mask, cnt = np.zeros(N, np.uint8), np.zeros(4, np.int32)
            best = lib.hc_recover_pose(vp(E), vp(m), N, ctypes.c_double(50.0), vp(Rg), vp(tg), vp(R), vp(t), vp(mask),
                                       vp(cnt), vp(err))
            assert best >= 0
            assert abs(R.reshape(3, 3) - g[f"R_{i}"][j].numpy()).max() < 1e-9
            assert abs(t - g[f"t_{i}"][j].numpy()).max() < 1e-9
            assert abs(err - g[f"err_{i}"][j].numpy()).max() < 1e-5
            if j == 0:
                assert (mask.astype(bool) == g[f"cv_mask_{i}"].numpy().astype(bool)).all()
                assert int(cnt.max()) == int(g[f"cv_n_{i}"])


def test_pose_loss_value_and_gradient_against_reference_autograd(lib, golden):
    """One PoseLoss term per model (Horn decomposition, cheirality vote, angular errors) and its gradient by
    forward-mode duals, against the reference's `PoseLoss.forward_average` + autograd run in fp64."""
    g = golden("pose_loss")
    for i in range(int(g["n_cases"])):
        m = np.ascontiguousarray(g[f"matches_{i}"].numpy(), dtype=np.float32)
        Rg = np.ascontiguousarray(g[f"R_gt_{i}"].numpy(), dtype=np.float64)
        tg = np.ascontiguousarray(g[f"t_gt_{i}"].numpy(), dtype=np.float64)
        M = g[f"E_{i}"].shape[0]
        total = 0.0
        for j in range(M):
            E = np.ascontiguousarray(g[f"E_{i}"][j].numpy(), dtype=np.float64)
            err, grad = np.zeros(2), np.zeros(9)
            assert lib.hc_pose_loss(vp(E), vp(m), m.shape[0], ctypes.c_double(50.0), vp(Rg), vp(tg), vp(err),
                                    vp(grad)) >= 0
            assert abs(err - g[f"err_{i}"][j].numpy()).max() < 1e-8
            want = g[f"grad_{i}"][j].numpy().ravel() * M          # forward_average divides by the number of models
            assert abs(grad - want).max() < 1e-7 * abs(want).max()
            total += err.mean()
        assert abs(total / M - float(g[f"loss_{i}"])) < 1e-9


# ---- tensor-core scorer (csrc/score_tc.cu, experimental): operand images, descriptors, column mapping ----
def _tc_scores(lib, matches, models, thr, words=2):
    m = np.ascontiguousarray(matches.numpy().astype(np.float32))
    md = np.ascontiguousarray(models.reshape(-1, 9).numpy().astype(np.float32))
    out = np.full(md.shape[0], -1.0, dtype=np.float32)
    lossless = ctypes.c_int(0)
    rc = lib.hc_msac_tc_scores(vp(m), m.shape[0], vp(md), md.shape[0], ctypes.c_float(thr), words, vp(out),
                               ctypes.byref(lossless))
    assert rc == 0, rc
    assert lossless.value == 1          # every operand word is an exact TF32: the hardware's truncation is a no-op
    return torch.from_numpy(out)


@pytest.mark.parametrize("words", [2, 3, 2 + 16, 3 + 16, 2 + 16 + 128, 3 + 16 + 128])
@pytest.mark.parametrize("N,K", [(2000, 40), (333, 30), (128, 20)])
def test_msac_tc_operands_reproduce_the_oracle_scores(lib, N, K, words):
    """The 3xTF32 contraction over 15 monomials, built and decoded exactly as score_tc.cu does it (images ->
    descriptors -> 128 x 256 x 8 MMA steps -> epilogue thread mapping), gives the soft-MSAC scores of
    msac_score.py:12-55 within the path's 1e-4 relative bar (ragged tails of points and of models included)."""
    from differentiable_ransac_b200 import synth
    from oracle import scoring

    matches, _, _ = synth.relative_pose_batch(1, N, seed=77, noise=5e-4)
    matches = matches[0]
    g = torch.Generator().manual_seed(N)
    idx = torch.stack([torch.randperm(N, generator=g)[:5] for _ in range(K)])
    inl = torch.arange(N - int(0.3 * N), N)
    idx[: K // 3] = inl[torch.stack([torch.randperm(len(inl), generator=g)[:5] for _ in range(K // 3)])]
    E = nister.five_point(matches[idx].double())
    E = E[trace_constraint_residual(E) < 1e-8].float()
    assert E.shape[0] > 128 or N < 2000   # more than one model tile at the headline size
    thr = 0.75 / 800.0
    want, _ = scoring.msac_score(matches.double(), E.double(), thr)
    got = _tc_scores(lib, matches, E, thr, words)
    rel = (got.double() - want).abs() / want.clamp_min(1.0)
    assert rel.max() < (1e-4 if words & 15 == 2 else 3e-5), rel.max()
    assert int(got.argmax()) == int(want.argmax())
    fp32, _ = scoring.msac_score(matches, E, thr)
    assert (got - fp32).abs().max() / fp32.max() < 1e-4


@pytest.mark.parametrize("words", [2, 3, 2 + 16, 3 + 16, 2 + 16 + 128, 3 + 16 + 128])
def test_msac_tc_nan_models_score_zero(lib, words):
    from differentiable_ransac_b200 import synth

    matches, E_gt, _ = synth.relative_pose_batch(1, 300, seed=5)
    models = torch.stack([E_gt[0], torch.full((3, 3), float("nan")), E_gt[0] / E_gt[0].norm()])
    got = _tc_scores(lib, matches[0], models, 0.75 / 800.0, words)
    # the NaN model scores 0 and -- pair reciprocal included -- leaves its column neighbour (model 0) alone;
    # model 2's neighbour is an absent model (odd count)
    assert got[1] == 0.0 and got[0] > 10 and abs(float(got[0] - got[2])) < 1e-3 * float(got[0])


def test_msac_tc_descriptors_match_the_cutlass_bitfields(lib, tmp_path):
    """smem / instruction descriptor words against cute::UMMA::SmemDescriptor / InstrDescriptor compiled on the
    host from the CUTLASS headers vendored in the image (skipped when they are absent)."""
    import glob
    import subprocess
    import sysconfig

    cands = glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "*", "data", "cutlass", "include")) + \
        glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "*", "3rdparty", "cutlass", "include"))
    cands = [c for c in cands if os.path.exists(os.path.join(c, "cute", "arch", "mma_sm100_desc.hpp"))]
    if not cands or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("no CUTLASS headers in this image")
    src = tmp_path / "desc_ref.cpp"
    src.write_text(r'''
#include <cstdint>
#include <cute/arch/mma_sm100_desc.hpp>
extern "C" uint64_t ref_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    cute::UMMA::SmemDescriptor d;
    d.desc_ = 0;
    d.version_ = 1;
    d.lbo_mode_ = 0;
    d.layout_type_ = uint8_t(cute::UMMA::LayoutType::SWIZZLE_NONE);
    d.start_address_ = uint16_t(addr >> 4);
    d.base_offset_ = 0;
    d.stride_byte_offset_ = sbo_bytes >> 4;
    d.leading_byte_offset_ = lbo_bytes >> 4;
    return d.desc_;
}
extern "C" uint32_t ref_instr_desc_tf32(int M, int N) {
    cute::UMMA::InstrDescriptor d;
    d.desc_ = 0;
    d.a_format_ = uint8_t(cute::UMMA::F16F32Format::TF32);
    d.b_format_ = uint8_t(cute::UMMA::F16F32Format::TF32);
    d.c_format_ = uint8_t(cute::UMMA::CFormat::F32);
    d.m_dim_ = M >> 4;
    d.n_dim_ = N >> 3;
    d.a_major_ = uint8_t(cute::UMMA::Major::K);
    d.b_major_ = uint8_t(cute::UMMA::Major::K);
    return d.desc_;
}
extern "C" uint32_t ref_instr_desc_bf16(int M, int N) {
    cute::UMMA::InstrDescriptor d;
    d.desc_ = 0;
    d.a_format_ = uint8_t(cute::UMMA::F16F32Format::BF16);
    d.b_format_ = uint8_t(cute::UMMA::F16F32Format::BF16);
    d.c_format_ = uint8_t(cute::UMMA::CFormat::F32);
    d.m_dim_ = M >> 4;
    d.n_dim_ = N >> 3;
    d.a_major_ = uint8_t(cute::UMMA::Major::K);
    d.b_major_ = uint8_t(cute::UMMA::Major::K);
    return d.desc_;
}
''')
    so = tmp_path / "desc_ref.so"
    subprocess.check_call(["g++", "-std=c++17", "-shared", "-fPIC", "-I", cands[0], "-I", "/usr/local/cuda/include",
                           str(src), "-o", str(so)])
    ref = ctypes.CDLL(str(so))
    ref.ref_smem_desc.restype = ctypes.c_uint64
    ref.ref_instr_desc_tf32.restype = ctypes.c_uint32
    lib.hc_tc_smem_desc.restype = ctypes.c_uint64
    lib.hc_tc_instr_desc.restype = ctypes.c_uint32
    for addr in (0x0, 0x400, 0x18000, 0x2fc00):
        assert lib.hc_tc_smem_desc(addr) == ref.ref_smem_desc(addr, 128, 1536)
    assert lib.hc_tc_instr_desc() == ref.ref_instr_desc_tf32(128, 256)
    lib.hc_tc_instr_desc_bf16.restype = ctypes.c_uint32
    ref.ref_instr_desc_bf16.restype = ctypes.c_uint32
    assert lib.hc_tc_instr_desc_bf16() == ref.ref_instr_desc_bf16(128, 256)
    assert lib.hc_tc_abytes() == 16 * 1536 and lib.hc_tc_bbytes() == 32 * 1536


def test_msac_tc_bf16_split_is_exact(lib):
    """Three BF16 words represent every fp32 exactly (the premise of the six-product variant)."""
    lib.hc_tc_bf16_sum.restype = ctypes.c_double
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.standard_normal(2000).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 4, 2000),
                         np.array([0.0, 1.0, -1.0, 1 / 3, 0.1, 1e-3, 123456.789], dtype=np.float32)])
    for x in xs.astype(np.float32):
        assert lib.hc_tc_bf16_sum(ctypes.c_float(float(x))) == float(x)


@pytest.mark.parametrize("words", [2, 3, 2 + 16])
@pytest.mark.parametrize("N,K", [(2000, 40), (333, 30), (80, 20), (90, 20)])
def test_msac_tc2_model_stationary_arrangement(lib, N, K, words):
    """Host model of csrc/score_tc2.cu: the models' words in tensor-memory order against correspondence tiles of
    80 rows read through the shared-memory descriptor (instruction shape 128 x 80), rows past N contributing 0."""
    from differentiable_ransac_b200 import synth
    from oracle import scoring

    matches, _, _ = synth.relative_pose_batch(1, N, seed=78, noise=5e-4)
    matches = matches[0]
    g = torch.Generator().manual_seed(N + 1)
    idx = torch.stack([torch.randperm(N, generator=g)[:5] for _ in range(K)])
    inl = torch.arange(N - int(0.3 * N), N)
    idx[: K // 3] = inl[torch.stack([torch.randperm(len(inl), generator=g)[:5] for _ in range(K // 3)])]
    E = nister.five_point(matches[idx].double())
    E = E[trace_constraint_residual(E) < 1e-8].float()
    thr = 0.75 / 800.0
    want, _ = scoring.msac_score(matches.double(), E.double(), thr)
    m = np.ascontiguousarray(matches.numpy().astype(np.float32))
    md = np.ascontiguousarray(E.reshape(-1, 9).numpy().astype(np.float32))
    out = np.full(md.shape[0], -1.0, dtype=np.float32)
    assert lib.hc_msac_tc2_scores(vp(m), N, vp(md), md.shape[0], ctypes.c_float(thr), words, vp(out)) == 0
    rel = (torch.from_numpy(out).double() - want).abs() / want.clamp_min(1.0)
    assert rel.max() < (1e-4 if words & 15 == 2 else 3e-5), rel.max()
    assert int(out.argmax()) == int(want.argmax())


@pytest.mark.parametrize("words", [2, 3, 3 + 16])
def test_msac_tc_against_the_reference_golden(lib, golden, words):
    """The tensor-core operand path (host model of csrc/score_tc.cu) against the scores the REFERENCE's own
    MSACScore.score produced (tests/golden/msac.npz, generated by importing scorings/msac_score.py): same
    scores within the path's 1e-4 relative bar, same arg-max."""
    g = golden("msac")
    got = _tc_scores(lib, g["matches"], g["models"], float(g["threshold"]), words)
    want = g["scores"]
    rel = (got - want).abs() / want.clamp_min(1.0)
    assert float(rel.max()) < 1e-4, float(rel.max())
    assert int(got.argmax()) == int(g["best"])


def test_e5_cooperative_stages_match_the_serial_solver_against_fp64():
    """csrc/e5_coop.cuh (four lanes per sample; on the host: four threads and a barrier) against the fp64 oracle
    (nister.py:69-408) beside the one-thread-per-sample composition of the same headers: the cooperative split
    changes summation orders and the pivoting bookkeeping, not what is found."""
    import ctypes

    import numpy as np

    import hostcheck
    from helpers import trace_constraint_residual
    from oracle import nister

    lib = hostcheck.load()
    from differentiable_ransac_b200 import synth

    m, _, inl = synth.relative_pose_pair(2000, 0.5, seed=3, noise=5e-4)
    g = torch.Generator().manual_seed(21)
    K = 256
    idx = torch.stack([torch.randperm(2000, generator=g)[:5].sort().values for _ in range(K)])
    inl_idx = inl.nonzero().flatten()
    for k in range(0, K, 2):                       # half the samples all-inlier
        idx[k] = inl_idx[torch.randperm(inl_idx.numel(), generator=g)[:5]].sort().values
    pts = m[idx]
    m64 = nister.five_point(pts.double()).reshape(K, 10, 3, 3)
    genuine = (trace_constraint_residual(m64.reshape(-1, 3, 3)) < 1e-8).reshape(K, 10)
    r = m64.flatten(2)
    r = r / r.norm(dim=-1, keepdim=True)
    ptsn = np.ascontiguousarray(pts.numpy(), dtype=np.float32)
    found = {}
    for fn, extra in (("hc_e5_solve_f32", False), ("hc_e5_coop_f32", True)):
        models = np.zeros((K, 10, 9), np.float32)
        nsol = np.zeros(K, np.int32)
        args = [ptsn.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(K), models.ctypes.data_as(ctypes.c_void_p),
                nsol.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(2)]
        if extra:
            args.append(ctypes.c_void_p(0))
        getattr(lib, fn)(*args)
        ours, ns = torch.from_numpy(models).reshape(K, 10, 3, 3).double(), torch.from_numpy(nsol)
        live = torch.arange(10)[None] < ns[:, None]
        assert torch.allclose(ours.flatten(2).norm(dim=-1)[live], torch.ones(int(live.sum()), dtype=torch.float64),
                              atol=1e-5)
        c = torch.where(live[..., None, None], ours, torch.full_like(ours, 1e3)).flatten(2)
        d = torch.minimum((c[:, :, None] - r[:, None]).norm(dim=-1), (c[:, :, None] + r[:, None]).norm(dim=-1)).min(1).values
        found[fn] = float((d[genuine] < 1e-3).double().mean())
        assert found[fn] >= 0.95 and float(d[genuine].median()) < 1e-5
    assert abs(found["hc_e5_solve_f32"] - found["hc_e5_coop_f32"]) < 0.01
