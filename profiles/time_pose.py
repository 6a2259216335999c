#!/usr/bin/env python
"""Time pose recovery / PoseLoss terms on one B200 (CUDA events): the evaluation of a test batch (one model per
pair), the GT-inlier masks of MatchLoss, and PoseLoss over K models per pair.  One JSON line each."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differentiable_ransac_b200 import ops, synth  # noqa: E402
from time_configs import timed  # noqa: E402

DEV = "cuda"


def main():
    B, N = 32, 2000
    data = [synth.relative_pose_pair(N, 0.5, seed=40 + b, noise=3e-4, return_pose=True) for b in range(B)]
    m = torch.stack([d[0] for d in data]).to(DEV)
    E = torch.stack([d[1] for d in data]).to(DEV)
    R = torch.stack([d[3] for d in data]).float().to(DEV)
    t = torch.stack([d[4] for d in data]).float().to(DEV)
    rows = [dict(stage="recover_pose + errors, 1 model per pair", M=1,
                 ms=timed(lambda: ops.recover_pose(E[:, None], m, None, R, t, want_mask=False))),
            dict(stage="GT-inlier masks (recover_pose with mask)", M=1, ms=timed(lambda: ops.recover_pose(E[:, None], m)))]
    for M in (64, 256):
        Es = E[:, None] + 0.01 * torch.randn(B, M, 3, 3, device=DEV)
        rows.append(dict(stage="pose_loss value + gradient", M=M, ms=timed(lambda: ops.pose_loss(Es, m, R, t), warm=2,
                                                                           reps=5)))
    for r in rows:
        r.update(B=B, N=N, triangulations_per_s=B * r["M"] * N * 4 / r["ms"] * 1e3)
        print(json.dumps(r))


if __name__ == "__main__":
    main()
