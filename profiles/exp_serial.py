"""Serial cfg2 steps: GPU time per step (events) with and without an L2 flush, and host time per step."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from differentiable_ransac_b200 import engine, ops
dev = torch.device("cuda", 0)
B, K, N = 32, 1000, 2000
mh, lh, th, _ = bench.make_inputs(B, N, seed=1234)
m, lg, thr = mh.to(dev), lh.to(dev), th.to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for i in range(5):
    engine.ransac_e5_test(m, lg, K, thr, seed=1, offset=i)
torch.cuda.synchronize()
for do_flush in (False, True):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    t0 = time.perf_counter()
    for i, (a, b) in enumerate(ev):
        if do_flush:
            flush.fill_(float(i))
        a.record()
        engine.ransac_e5_test(m, lg, K, thr, seed=1, offset=10 + i)
        b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print(json.dumps(dict(kernel=ops._MSAC_KERNEL, flush=do_flush, ms_median=ts[25], ms_min=ts[0], ms_max=ts[-1],
                          host_issue_ms_per_step=(t1 - t0) * 1e3 / 50)), flush=True)
