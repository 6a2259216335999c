#!/usr/bin/env python
"""Time the post-loop stages at the headline shape on one B200 with CUDA events:
the non-minimal fits alone, one LO iteration (fit on inliers + score + winner mask) and the final refit
appended to the cfg2 test-mode step.  One JSON line per measurement."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differentiable_ransac_b200 import engine, ops, synth  # noqa: E402
from time_configs import timed  # noqa: E402

DEV = "cuda"


def main():
    B, K, N = 32, 1000, 2000
    m, E_gt, _ = synth.relative_pose_batch(B, N, noise=3e-4)
    m = m.to(DEV)
    lg = synth.logits_regime(B, N, "L0", seed=1).to(DEV)
    thr = torch.full((B,), 0.75 / 800.0, device=DEV)
    it = [0]

    def loop():
        it[0] += 1
        return engine.ransac_e5_test(m, lg, K, thr, seed=5, offset=it[0])

    state = loop()
    rows = []
    rows.append(dict(stage="refit_e5 all points", ms=timed(lambda: ops.refit_e5(m, None))))
    rows.append(dict(stage="refit_e5 on inlier masks", ms=timed(lambda: ops.refit_e5(m, state["mask"]))))
    rows.append(dict(stage="refit_f8 on inlier masks", ms=timed(lambda: ops.refit_f8(m, state["mask"]))))
    rows.append(dict(stage="one LO iteration (fit + score + mask)",
                     ms=timed(lambda: engine._fit_and_score(m, state["mask"], thr, False))))
    rows.append(dict(stage="cfg2 loop only", ms=timed(loop)))
    rows.append(dict(stage="cfg2 loop + final refit", ms=timed(lambda: engine.final_refit(m, loop(), thr, False))))
    for r in rows:
        r.update(B=B, K=K, N=N)
        print(json.dumps(r))


if __name__ == "__main__":
    main()
