"""world_size-2 gloo tests of the multi-GPU plumbing (pair sharding, gradient all-reduce,
result gather) on the CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from differentiable_ransac_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(0)
        data = torch.randn(B, 6, generator=gen)
        target = torch.randn(B, 1, generator=gen)
        (local, tgt) = parallel.shard_pairs(data, target)
        lo, hi = parallel.shard_range(B, rank, world)
        assert local.shape[0] == hi - lo and torch.equal(local, data[lo:hi])
        # a toy "weight network": per-pair losses, summed locally, gradients summed over ranks
        torch.manual_seed(1)
        net = torch.nn.Linear(6, 1)
        loss = ((net(local) - tgt) ** 2).sum()
        loss.backward()
        parallel.allreduce_gradients(list(net.parameters()), average=False, bucket_bytes=8)
        torch.manual_seed(1)
        ref = torch.nn.Linear(6, 1)
        ((ref(data) - target) ** 2).sum().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-5), (p.grad, q.grad)
        # gather per-pair results in rank order
        res = parallel.gather_results(local * 2.0)
        assert torch.equal(res, data * 2.0)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_allreduce_gather():
    mp.spawn(_worker, args=(2, _free_port(), 7), nprocs=2, join=True)
