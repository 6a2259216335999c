"""Round-2 first measurement of the EXPERIMENTAL tensor-core scorer (csrc/score_tc.cu) against the two FP32
kernels on the cfg2 compact model list (B200; CUDA events, L2 flushed).  Run under a timeout -- the kernel has
never run on hardware and a wrong mbarrier phase is a hang:

    timeout 120 python profiles/time_score_tc.py [B]

Prints one JSON line per kernel (median / min ms, models scored) and the largest relative difference of the
"tc" scores from the "block" scores."""
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
m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)

KERNS = tuple(sys.argv[2:]) or ("block", "stream", "tc_tf32", "tc_bf16", "tc_tf32p", "tc_bf16p", "tc_tf32q", "tc_bf16q", "tc_tf32_e16", "tc_tf32p_e16", "tc_bf16p_e16",
             "tc2_tf32", "tc2_tf32_e16", "tc2_bf16", "tc2_tf32p", "tc2_tf32p_e16")
s_ref, b_ref = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel="block")
s_tc, b_tc = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel=KERNS[-1] if sys.argv[2:] else "tc")
torch.cuda.synchronize()
live = torch.arange(cm.shape[1], device=dev)[None] < cc[:, None]
rel = ((s_tc - s_ref).abs() / s_ref.clamp_min(1.0))[live]
print(json.dumps(dict(check="tc vs block", max_rel=float(rel.max()), same_best=int((b_tc == b_ref).sum()), pairs=B)),
      flush=True)

for kern in KERNS * 2:
    for _ in range(3):
        ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, kernel=kern)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        flush.fill_(1.0)
        best = torch.zeros(B, dtype=torch.int64, device=dev)
        a.record()
        ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, best=best, kernel=kern)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print(json.dumps(dict(B=B, kernel=kern, ms_median=ts[len(ts) // 2], ms_min=ts[0], models=int(cc.sum()))), flush=True)
