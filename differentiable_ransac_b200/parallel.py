"""Multi-GPU plumbing: one process per GPU, image pairs sharded over ranks.

The hypothesize-and-score path needs NO collective in the forward pass -- every pair is
independent (`model_cl.py:488` carries no cross-pair state, RANSACLayer has no parameters).
Training adds one all-reduce of the weight-network's parameter gradients; d loss / d logits is
per pair and is consumed locally by the network's backward.  torch.distributed (NCCL on the GPUs,
gloo in the CPU tests) is the transport."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of `n_items` pairs for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs(*tensors, rank=None, world=None):
    """Slice the leading (pair) dimension of every tensor for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return tuple(t[lo:hi] for t in tensors)


def allreduce_gradients(params, group=None, average=True, bucket_bytes=4 << 20):
    """Data-parallel gradient all-reduce (sum or mean) in flat buckets: 2.5 MB of CLNet gradients is
    latency-bound on NVLink, so few large messages beat one per tensor."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def gather_results(local: torch.Tensor, group=None):
    """Concatenate per-rank results (e.g. best models [B_local,3,3]) on every rank, in rank order.
    Shards may differ in length by one."""
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: int(s)] for o, s in zip(out, sizes)])
