#!/usr/bin/env python
"""Time BASELINE.json configs 3-5 (full size) on one B200 with CUDA events; prints one JSON line per
config.  Parity for these shapes is in tests/test_gpu_full_size.py; the headline cfg2 is bench.py."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differentiable_ransac_b200 import engine, synth  # noqa: E402

DEV = "cuda"


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def cfg3():
    B, K, N = 64, 2000, 2000
    ms, inls = zip(*[(lambda r: (r[0] / 640.0, r[3]))(synth.pixel_pair(N, 0.5, seed=100 + b)) for b in range(B)])
    matches = torch.stack(ms).to(DEV)
    logits = synth.logits_regime(B, N, "L1", seed=1).to(DEV)
    P = max(int(i.sum()) for i in inls)
    pts = torch.zeros(B, P, 4)
    npts = torch.tensor([int(i.sum()) for i in inls], dtype=torch.int32)
    for b in range(B):
        pts[b, : int(npts[b])] = ms[b][inls[b]]
    pts, npts = pts.to(DEV), npts.to(DEV)
    it = [0]

    def step():
        m = matches.clone().requires_grad_(True)
        lg = logits.clone().requires_grad_(True)
        models, valid = engine.HypothesizeF8.apply(m, lg, K, 1.0, None, 9, it[0])
        engine.match_loss(models, valid, pts, npts).mean().backward()
        it[0] += 1

    ms_ = timed(step)
    return dict(config="cfg3 F-8PC train fwd+bwd", B=B, K=K, N=N, ms_per_step=ms_, hyps_per_s=B * K / ms_ * 1e3)


def cfg4():
    B, K, N = 16, 1000, 50000
    pts = torch.stack([synth.rigid_pair(N, 0.7, seed=200 + b)[0] for b in range(B)]).to(DEV)
    logits = synth.logits_regime(B, N, "L1", seed=2).to(DEV)
    it = [0]

    def fwd():
        models, valid = engine.HypothesizeRigid.apply(pts, logits, K, True, 1.0, None, 3, it[0])
        engine.RigidResidual.apply(pts, models)
        it[0] += 1

    def fwdbwd():
        lg = logits.clone().requires_grad_(True)
        models, valid = engine.HypothesizeRigid.apply(pts, lg, K, True, 1.0, None, 3, it[0])
        engine.RigidResidual.apply(pts, models).mean().backward()
        it[0] += 1

    a, b = timed(fwd), timed(fwdbwd)
    return dict(config="cfg4 rigid 3-pt train", B=B, K=K, N=N, ms_fwd=a, ms_fwd_bwd=b, hyps_per_s_fwd=B * K / a * 1e3,
                hyps_per_s_fwd_bwd=B * K / b * 1e3)


def cfg5(K=1000):
    B, N = 32, 2000
    matches, E_gt, inl = synth.relative_pose_batch(B, N, seed=300, noise=2e-4)
    logits = synth.logits_regime(B, N, "L0", seed=4).to(DEV)
    P = int(inl.sum(1).max())
    pts = torch.zeros(B, P, 4)
    npts = inl.sum(1).int()
    for b in range(B):
        pts[b, : int(npts[b])] = matches[b][inl[b]]
    matches, E_gt, pts, npts = matches.to(DEV), E_gt.to(DEV), pts.to(DEV), npts.to(DEV)
    it = [0]

    def step():
        m = matches.clone().requires_grad_(True)
        lg = logits.clone().requires_grad_(True)
        chosen, valid = engine.HypothesizeE5.apply(m, lg, E_gt, K, 1.0, None, 5, it[0], True)
        engine.match_loss(chosen, valid, pts, npts).mean().backward()
        it[0] += 1

    ms_ = timed(step)
    return dict(config=f"cfg5 5PC train fwd+bwd, 32 pairs/GPU, K={K}", B=B, K=K, N=N, ms_per_step=ms_,
                hyps_per_s=B * K / ms_ * 1e3)


if __name__ == "__main__":
    for f in (cfg3, cfg4, cfg5, lambda: cfg5(128)):
        print(json.dumps(f()))
