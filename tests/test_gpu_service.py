"""GPU tests of engine.E5TestService: the pipelined host-buffer entry of the test-mode path returns, batch
for batch, exactly what the direct device call returns for the same (seed, offset), whatever the number of
batches in flight."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("slots", [1, 2, 3])
def test_service_matches_direct_call(slots, graph):
    from differentiable_ransac_b200 import engine, synth

    B, N, K = 6, 700, 96
    batches = []
    for j in range(5):
        matches, _, _ = synth.relative_pose_batch(B, N, seed=40 + j, noise=5e-4)
        logits = synth.logits_regime(B, N, "L0", seed=50 + j)
        thr = torch.full((B,), 0.75 / 800.0)
        batches.append((matches, logits, thr))
    svc = engine.E5TestService(B, N, K, DEV, slots=slots, seed=11, graph=graph)
    pending, got = [], []
    for matches, logits, thr in batches:
        if len(pending) == slots:
            got.append({k: v.clone() for k, v in svc.result(pending.pop(0)).items()})
        slot = svc.step % slots
        svc.stage(slot, matches, logits, thr)
        pending.append(svc.submit(slot))
    for s in pending:
        got.append({k: v.clone() for k, v in svc.result(s).items()})
    svc.drain()
    assert len(got) == len(batches)
    for j, (matches, logits, thr) in enumerate(batches):
        want = engine.ransac_e5_test(matches.to(DEV), logits.to(DEV), K, thr.to(DEV), seed=11, offset=j,
                                     scorer=svc.scorer)   # the kernel the service picks
        torch.cuda.synchronize()
        assert torch.equal(got[j]["best_id"], want["best_id"].cpu())
        assert torch.equal(got[j]["best_score"], want["best_score"].cpu())
        assert torch.equal(got[j]["best_model"], want["best_model"].cpu())
        assert torch.equal(got[j]["ninl"], want["ninl"].cpu())


@pytest.mark.parametrize("graph", [False, True])
def test_device_resident_service_matches_direct_call(graph):
    from differentiable_ransac_b200 import engine, synth

    B, N, K, slots = 4, 600, 64, 2
    svc = engine.E5TestService(B, N, K, DEV, slots=slots, seed=5, graph=graph, host_io=False)
    batches = []
    for j in range(4):
        matches, _, _ = synth.relative_pose_batch(B, N, seed=70 + j, noise=5e-4)
        logits = synth.logits_regime(B, N, "L0", seed=80 + j)
        thr = torch.full((B,), 0.75 / 800.0)
        batches.append((matches.to(DEV), logits.to(DEV), thr.to(DEV)))
    for j, (m, lg, thr) in enumerate(batches):
        slot = svc.submit(packed=torch.cat((m.flatten(), lg.flatten(), thr)))
        got = {k: v.clone() for k, v in svc.result(slot).items()}
        want = engine.ransac_e5_test(m, lg, K, thr, seed=5, offset=j, scorer=svc.scorer)
        torch.cuda.synchronize()
        assert torch.equal(got["best_id"], want["best_id"])
        assert torch.equal(got["best_score"], want["best_score"])
        assert torch.equal(got["best_model"], want["best_model"])
