"""Hardware check of the tensor-core scorer (csrc/score_tc.cu), both operand splits: parity cases of growing size
(ragged counts, empty pairs, more pairs than a warp) against the FP32 block kernel, then the cfg2 timing.  Every
line is flushed (and appended to gpurun_out/tc_try.jsonl) so a hang shows how far it got.  Run under `timeout`."""
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


def ids_of(best):
    return [(0xFFFFFFFF - (int(k) & 0xFFFFFFFF)) if int(k) else -1 for k in best.cpu()]


from differentiable_ransac_b200 import ops, synth  # noqa: E402

dev = "cuda"
KERNELS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tc_tf32", "tc_bf16", "tc_tf32p", "tc_bf16p"]
say(stage="start", t=time.time(), kernels=KERNELS)
CASES = [(1, 5, 3, None), (2, 33, 64, None), (3, 70, 257, [70, 0, 41]), (3, 300, 2500, [300, 0, 129]),
         (4, 1000, 2000, [1000, 517, 1, 32]), (40, 200, 500, None)]
for kern in KERNELS:
    for (B, M, N, counts) in CASES:
        matches, _, _ = synth.relative_pose_batch(B, max(N, 8), seed=5 + N)
        matches = matches[:, :N].contiguous().to(dev)
        gen = torch.Generator().manual_seed(M)
        models = torch.randn(B, M, 3, 3, generator=gen)
        models = (models / models.flatten(-2).norm(dim=-1)[..., None, None]).to(dev)
        thr = (torch.rand(B, generator=gen) * 0.05 + 0.002).to(dev)
        count = None if counts is None else torch.tensor(counts, dtype=torch.int32, device=dev)
        ids = torch.stack([torch.randperm(4 * M + 7, generator=gen)[:M] for _ in range(B)]).int().to(dev)
        s_ref, b_ref = ops.score_msac(matches, models, thr, count=count, ids=ids, kernel="block")
        torch.cuda.synchronize()
        s_tc, b_tc = ops.score_msac(matches, models, thr, count=count, ids=ids, kernel=kern)
        s_2, b_2 = ops.score_msac(matches, models, thr, count=count, ids=ids, kernel=kern)
        torch.cuda.synchronize()
        cnt = torch.full((B,), M, device=dev) if count is None else count
        live = torch.arange(M, device=dev)[None, :] < cnt[:, None]          # [B, M]
        rel = ((s_tc - s_ref).abs() / s_ref.clamp_min(1.0))[live]
        say(stage="parity", kernel=kern, case=[B, M, N, counts], max_rel=float(rel.max()),
            same_ids=ids_of(b_tc) == ids_of(b_ref), deterministic=bool(torch.equal(s_tc[live], s_2[live]) and torch.equal(b_tc, b_2)))

import bench  # noqa: E402

B, K, N = 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
models, nsol, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
s_ref, b_ref = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel="block")
live = torch.arange(cm.shape[1], device=dev)[None] < cc[:, None]
for kern in KERNELS:
    s_tc, b_tc = ops.score_msac(m, cm, thr, count=cc, ids=cid, want_scores=True, kernel=kern)
    torch.cuda.synchronize()
    rel = ((s_tc - s_ref).abs() / s_ref.clamp_min(1.0))[live]
    say(stage="parity", kernel=kern, case="cfg2", max_rel=float(rel.max()), mean_rel=float(rel.mean()),
        same_ids=sum(int(a == b) for a, b in zip(ids_of(b_tc), ids_of(b_ref))), pairs=B)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for kern in ["block", "stream"] + KERNELS:
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
