"""solve_e5 at the cfg2 shape (32 pairs x 1000 hypotheses, minimal samples drawn by drb_sample_sets), CUDA events,
L2 flushed before every launch.  DRB_E5_SOLVER=thread selects the round-1 one-thread-per-hypothesis kernel
(read once per process), the default is the cooperative four-lanes-per-hypothesis kernel (csrc/e5_coop.cuh).

    python profiles/time_solve_e5.py [B]           # prints one JSON line
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from differentiable_ransac_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
m, lg = matches_h.to(dev), logits_h.to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
for _ in range(3):
    models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for a, b in ev:
    flush.fill_(1.0)
    cc0 = torch.zeros(B, dtype=torch.int32, device=dev)
    a.record()
    ops.solve_e5(m, idx, compact=True, ccount=cc0)
    b.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
print(json.dumps(dict(kernel=os.environ.get("DRB_E5_SOLVER", "coop"), B=B, ms_median=ts[len(ts) // 2], ms_min=ts[0],
                      models=int(cc.sum()), nsol_mean=float(nsol.float().mean()))))
