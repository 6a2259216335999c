"""First contact of the experimental tensor-core scorer with hardware: three parity cases of growing size against
the FP32 block kernel, then the cfg2 timing.  Every line is flushed (and appended to gpurun_out/tc_try.jsonl) so a
hang shows how far it got.  Run under `timeout`."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "tc_try.jsonl"), "a")


def say(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


from differentiable_ransac_b200 import ops, synth  # noqa: E402

dev = "cuda"
say(stage="start", t=time.time())
for (B, M, N) in ((1, 5, 3), (2, 33, 64), (3, 300, 2500)):
    matches, _, _ = synth.relative_pose_batch(B, max(N, 8), seed=5 + N)
    matches = matches[:, :N].contiguous().to(dev)
    gen = torch.Generator().manual_seed(M)
    models = torch.randn(B, M, 3, 3, generator=gen)
    models = (models / models.flatten(-2).norm(dim=-1)[..., None, None]).to(dev)
    thr = (torch.rand(B, generator=gen) * 0.05 + 0.002).to(dev)
    s_ref, b_ref = ops.score_msac(matches, models, thr, kernel="block")
    torch.cuda.synchronize()
    say(stage="launch tc", case=[B, M, N])
    s_tc, b_tc = ops.score_msac(matches, models, thr, kernel="tc")
    torch.cuda.synchronize()
    rel = (s_tc - s_ref).abs() / s_ref.clamp_min(1.0)
    say(stage="done", case=[B, M, N], max_rel=float(rel.max()), same_best=bool((b_tc == b_ref).all()),
        tc_head=[float(x) for x in s_tc.flatten()[:4]], ref_head=[float(x) for x in s_ref.flatten()[:4]])

import bench  # noqa: E402

B, K, N = 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
s_ref, b_ref = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel="block")
torch.cuda.synchronize()
say(stage="launch tc", case="cfg2")
s_tc, b_tc = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel="tc")
torch.cuda.synchronize()
live = torch.arange(cm.shape[1], device=dev)[None] < cc[:, None]
rel = ((s_tc - s_ref).abs() / s_ref.clamp_min(1.0))[live]
say(stage="done", case="cfg2", max_rel=float(rel.max()), same_best=int((b_tc == b_ref).sum()), pairs=B)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for kern in ("block", "stream", "tc"):
    for _ in range(3):
        ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, kernel=kern)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        flush.fill_(1.0)
        best = torch.zeros(B, dtype=torch.int64, device=dev)
        a.record()
        ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=False, best=best, kernel=kern)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    say(stage="time", kernel=kern, ms_median=ts[len(ts) // 2], ms_min=ts[0], models=int(cc.sum()))
