"""GPU tests of the device-side replay of the reference's chunked test-mode loop with adaptive termination
(`drb_adaptive_select`).  End-to-end parity with the reference's own `RANSAC.__call__` is in
tests/test_gpu_refit.py (driver_full_* fixtures run through this path when lo = 0); here the replay is checked
against the loop itself, chunk by chunk, on the same hypotheses."""
import math
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _reference_loop(run, m, lg, thr, noise, rbs, max_iterations, sample_size, confidence=0.999, eps=1e-5):
    """ransac.py:55-144 spelled out on the host for ONE pair over per-chunk launches (one sync per chunk)."""
    N = m.shape[1]
    best, iterations, max_iters, c = None, 0, max_iterations, 0
    while iterations < max_iters:
        out = run(m, lg, rbs, thr, noise=noise[:, c * rbs:(c + 1) * rbs])
        if best is None or bool(out["best_score"][0] > best["best_score"][0]):
            best = out
            ratio = int(out["ninl"][0]) / N
            if 1.0 - ratio ** sample_size >= 1.0 - eps:
                a = max_iterations
            else:
                a = max(0.0, math.log10(1.0 - confidence) / math.log10(1 - ratio ** sample_size + eps))
            max_iters = min(max_iterations, a)
        iterations += rbs
        c += 1
    return best, iterations


@pytest.mark.parametrize("sample_size", [5, 8])
def test_replay_equals_the_loop_pair_by_pair(sample_size):
    from differentiable_ransac_b200 import engine, synth
    B, N, rbs, max_it = 6, 1000, 32, 300                 # 300 is not a multiple of 32: 10 chunks, like the reference
    C = -(-max_it // rbs)
    ratios = (0.9, 0.8, 0.6, 0.5, 0.3, 0.95)
    if sample_size == 5:
        pairs = [synth.relative_pose_pair(N, ratios[b], seed=500 + b, noise=2e-4) for b in range(B)]
        m = torch.stack([p[0] for p in pairs]).to(DEV)
        thr = torch.full((B,), 0.75 / 800.0, device=DEV)
        run = engine.ransac_e5_test
    else:
        pairs = [synth.pixel_pair(N, ratios[b], seed=600 + b) for b in range(B)]
        m = torch.stack([p[0] for p in pairs]).to(DEV)
        thr = torch.full((B,), 0.75, device=DEV)
        run = engine.ransac_f8_test
    lg = synth.logits_regime(B, N, "L1", seed=4).to(DEV)
    noise = synth.gumbel_noise((B, C * rbs, N), seed=77).to(DEV)
    out = engine.ransac_test_adaptive(m, lg, rbs, max_it, thr, sample_size, noise=noise)
    seen = set()
    for b in range(B):
        ref, its = _reference_loop(run, m[b:b + 1], lg[b:b + 1], thr[b:b + 1], noise[b:b + 1], rbs, max_it, sample_size)
        assert int(out["iterations"][b]) == its
        assert float(out["best_score"][b]) == float(ref["best_score"][0])
        assert torch.equal(out["best_model"][b], ref["best_model"][0])
        assert torch.equal(out["mask"][b], ref["mask"][0]) and int(out["ninl"][b]) == int(ref["ninl"][0])
        seen.add(its)
    assert len(seen) > 1 and min(seen) < C * rbs          # the pairs really stop at different trip counts


def test_adaptive_exponent_is_the_estimators_sample_size():
    """`-fmat 1 -sam 3`: the sampler draws 8 points but the iteration budget uses `estimator.sample_size` = 7
    (ransac.py:207,214; fundamental_matrix_estimator.py:163).  The replay with exponent 7 equals the loop run with
    exponent 7, and differs from exponent 8 where the two budgets fall into different chunks."""
    from differentiable_ransac_b200 import engine, synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    B, N, rbs, max_it = 4, 1000, 32, 640
    ratios = (0.62, 0.55, 0.5, 0.7)
    pairs = [synth.pixel_pair(N, ratios[b], seed=650 + b) for b in range(B)]
    m = torch.stack([p[0] for p in pairs]).to(DEV)
    thr = torch.full((B,), 0.75, device=DEV)
    lg = synth.logits_regime(B, N, "L1", seed=4).to(DEV)
    noise = synth.gumbel_noise((B, max_it, N), seed=78).to(DEV)
    out7 = engine.ransac_test_adaptive(m, lg, rbs, max_it, thr, 8, noise=noise, adaptive_exponent=7)
    out8 = engine.ransac_test_adaptive(m, lg, rbs, max_it, thr, 8, noise=noise)
    for b in range(B):
        _, its7 = _reference_loop(engine.ransac_f8_test, m[b:b + 1], lg[b:b + 1], thr[b:b + 1], noise[b:b + 1], rbs,
                                  max_it, 7)
        assert int(out7["iterations"][b]) == its7
    assert (out7["iterations"] <= out8["iterations"]).all() and (out7["iterations"] < out8["iterations"]).any()
    # and the driver passes the estimator's size, whatever the sampler draws
    opt = types.SimpleNamespace(device=DEV, fmat=1, sampler=3, precision=1, tr=0, threshold=0.75,
                                ransac_batch_size=rbs, weighted=0)
    drv = RANSACLayer(opt).estimator
    drv.max_iterations = max_it
    assert drv.sampler.num_samples == 8 and drv.estimator.sample_size == 7
    got = drv._loop(m, lg, thr, noise)
    assert torch.equal(got["iterations"], out7["iterations"])


def test_seven_point_and_philox_paths_run_and_stop_early():
    from differentiable_ransac_b200 import engine, synth
    B, N = 4, 1500
    pairs = [synth.pixel_pair(N, 0.9, seed=700 + b) for b in range(B)]
    m = torch.stack([p[0] for p in pairs]).to(DEV)
    lg = synth.logits_regime(B, N, "L1", seed=5).to(DEV)
    out = engine.ransac_test_adaptive(m, lg, 64, 2048, torch.full((B,), 0.75, device=DEV), 7, seed=3)
    assert (out["iterations"] < 2048).all() and (out["iterations"] % 64 == 0).all()
    assert (out["ninl"] > 0.8 * N).all()


def test_driver_and_batched_entry_agree_on_iterations(golden):
    """`RANSAC.__call__` per pair and `batched_test` for the batch return the same winners and trip counts."""
    from differentiable_ransac_b200 import synth
    from differentiable_ransac_b200.model_cl import RANSACLayer
    B, N = 4, 1000
    pairs = [synth.relative_pose_pair(N, (0.9, 0.7, 0.5, 0.3)[b], seed=800 + b, noise=2e-4) for b in range(B)]
    m = torch.stack([p[0] for p in pairs]).to(DEV)
    lg = synth.logits_regime(B, N, "L0", seed=6).to(DEV)
    K1 = torch.tensor([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    opt = types.SimpleNamespace(device=DEV, fmat=0, sampler=2, precision=1, tr=0, threshold=0.75,
                                ransac_batch_size=64, weighted=0)
    drv = RANSACLayer(opt).estimator
    drv.max_iterations = 1024
    noise = synth.gumbel_noise((B, 1024, N), seed=9).to(DEV)
    per_pair = []
    for b in range(B):
        drv.sampler.injected_noise = noise[b]
        per_pair.append(drv(m[b], lg[b], K1, K1, None))
    drv.sampler.injected_noise = None
    thr = torch.full((B,), 0.75 / 800.0, device=DEV)
    out = drv._loop(m, lg, thr, noise)
    for b, (model, mask, score, its) in enumerate(per_pair):
        assert int(out["iterations"][b]) == its
        assert torch.equal(out["mask"][b], mask)
    assert per_pair[0][3] < per_pair[3][3]               # 90 % inliers stops long before 30 %
