"""Two launches of the tensor-core scorer at the cfg2 shape, for `ncu -k regex:score_msac_tc_kernel -s 1 -c 1`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from differentiable_ransac_b200 import ops  # noqa: E402

dev = "cuda"
kern = sys.argv[1] if len(sys.argv) > 1 else "tc"
B, K, N = 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)     # twice: `ncu -k regex:solve_e5_kernel -s 1 -c 1` takes the warm one
for _ in range(2):
    ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, kernel=kern)
torch.cuda.synchronize()
print("ok")
