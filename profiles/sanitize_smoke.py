#!/usr/bin/env python
"""One small call of every kernel, for `compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_smoke.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differentiable_ransac_b200 import engine, ops, synth  # noqa: E402

DEV = "cuda"
B, K, N = 3, 37, 333                      # awkward sizes on purpose: N % 4 != 0, K % 32 != 0
m, E, inl = synth.relative_pose_batch(B, N, seed=1, noise=2e-4)
lg = synth.logits_regime(B, N, "L1", seed=2)
thr = torch.full((B,), 1e-3)
noise = synth.gumbel_noise((B, K, N), seed=3)
m, E, lg, thr, noise = m.to(DEV), E.to(DEV), lg.to(DEV), thr.to(DEV), noise.to(DEV)
for kw in (dict(), dict(noise=noise), dict(sampler="gumbel")):
    out = engine.ransac_e5_test(m, lg, K, thr, want_scores=True, **kw)
TC = os.environ.get("DRB_SANITIZE_TC", "tc_tf32,tc_bf16,tc_bf16p,tc_bf16p_s,tc_bf16q,tc_tf32q").split(",")
for scorer in ["block", "stream"] + [t for t in TC if t]:   # every MSAC kernel; the queue kernel with a block cut into pieces
    out = engine.ransac_e5_test(m, lg, K, thr, want_scores=True, scorer=scorer)
svc = engine.E5TestService(B, N, K, DEV, slots=2, seed=1, graph=False, host_io=False)
for _ in range(3):
    svc.submit(packed=torch.cat((m.flatten(), lg.flatten(), thr)))
svc.drain()
out8 = engine.ransac_f8_test(m, lg, K, thr)
ops.solve_f7(m, ops.sample_sets(lg, K, 7))
mm = m.clone().requires_grad_(True)
ll = lg.clone().requires_grad_(True)
ch, v = engine.HypothesizeE5.apply(mm, ll, E, K, 1.0, None, 1, 2, True)
engine.match_loss(ch, v, m).mean().backward()
mm.grad = None
f, v = engine.HypothesizeF8.apply(mm, ll, K, 0.7, noise, 0, 0)
engine.match_loss(f, v, m, torch.tensor([N, N - 5, 7], dtype=torch.int32, device=DEV)).mean().backward()
pts = torch.stack([synth.rigid_pair(1001, 0.5, seed=s)[0] for s in range(2)]).to(DEV)   # odd N: non-bulk tail tile
l3 = torch.rand(2, 1001, device=DEV, requires_grad=True)
for flag in (True, False):
    r, v = engine.HypothesizeRigid.apply(pts, l3, 50, flag, 1.0, None, 0, 0)
    engine.RigidResidual.apply(pts, r).mean().backward()
for kind, data in (("e5", (m, lg, E, m, torch.full((B,), N, dtype=torch.int32, device=DEV))),
                   ("f8", (m, lg, None, m, torch.full((B,), N, dtype=torch.int32, device=DEV)))):
    st = engine.TrainStep(kind, B, N, K, DEV, P=N, seed=3, graph=False, want_grad_matches=True)   # the fused training step
    st.run(*data)
st = engine.TrainStep("rigid", 2, 1001, 50, DEV, seed=3, graph=False)
st.run(pts, l3.detach())
out64 = engine.ransac_e5_test_f64(m, lg, K, thr, noise=noise, want_scores=True)      # the float64 chain (fp64_path.cu)
ops.rigid_residual_forward_backward(pts, r.detach(), torch.ones(2, 50, device=DEV))  # moments kernel, N = 1001
ops.rigid_residual_forward(pts, r.detach(), want_ninl=True)                          # per-point kernel (inlier counts)
ops.rigid_residual_forward(pts, r.detach(), want_ninl=False)                         # moments kernel
torch.cuda.synchronize()
print("sanitize smoke ok", float(out["best_score"].sum()), float(out8["best_score"].sum()))
