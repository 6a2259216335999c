"""Experiment (GPU box): can the latency-bound 5-point kernel hide behind the FMA-bound scorer?

Variants of the cfg2 step (32 pairs x 1000 hyps x 2000 corrs), all timed with CUDA events, L2 flushed
between steps:
  eager1      the product path as it is (one stream, 4 launches)
  graph1      the same, replayed from a CUDA graph (how much of the step is CPU launch overhead?)
  stagC       pairs split into C chunks; a solver stream (high priority) runs sample+solve for chunk c
              while the scoring stream scores chunk c-1 (software pipeline), eager
  gstagC      the same captured in a CUDA graph
Prints one JSON line per variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from differentiable_ransac_b200 import engine, ops  # noqa: E402

dev = torch.device("cuda", 0)
B, K, N = 32, 1000, 2000
matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
matches, logits, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def time_it(fn, steps=30, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(float(i))
        ev[i][0].record()
        fn(warm + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return dict(ms_mean=sum(ts) / len(ts), ms_median=ts[len(ts) // 2], ms_min=ts[0])


def eager1(i):
    return engine.ransac_e5_test(matches, logits, K, thr, seed=42, offset=i)


lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
s_solve = torch.cuda.Stream(device=dev, priority=-1)
s_score = torch.cuda.Stream(device=dev, priority=0)


def staggered(C, i, prio=True):
    main = torch.cuda.current_stream()
    start = torch.cuda.Event()
    start.record(main)
    ss = s_solve if prio else s_score2
    ss.wait_event(start)
    s_score.wait_event(start)
    outs = []
    solved = []
    with torch.cuda.stream(ss):
        for c in range(C):
            b0, b1 = B * c // C, B * (c + 1) // C
            idx = ops.sample_sets(logits[b0:b1], K, 5, 42, i + (c << 32))
            best0, cc0 = ops.zeroed_counters(b1 - b0, dev)
            models, nsol, cm, cid, cc = ops.solve_e5(matches[b0:b1], idx, compact=True, ccount=cc0)
            e = torch.cuda.Event()
            e.record(ss)
            solved.append((b0, b1, models, cm, cid, cc, best0, e))
    with torch.cuda.stream(s_score):
        for (b0, b1, models, cm, cid, cc, best0, e) in solved:
            s_score.wait_event(e)
            _, best = ops.score_msac(matches[b0:b1], cm, thr[b0:b1], count=cc, ids=cid, want_scores=False, best=best0)
            outs.append(ops.best_finalize(matches[b0:b1], models.reshape(b1 - b0, -1, 9), best, thr[b0:b1]))
    main.wait_stream(ss)
    main.wait_stream(s_score)
    return outs


s_score2 = torch.cuda.Stream(device=dev, priority=0)


def graphed(fn):
    # warm up on a side stream, then capture
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn(7)
    return g, out


res = {}
res["eager1"] = time_it(eager1)
print(json.dumps({"variant": "eager1", **res["eager1"]}), flush=True)
try:
    g, _ = graphed(eager1)
    r = time_it(lambda i: g.replay())
    print(json.dumps({"variant": "graph1", **r}), flush=True)
except Exception as ex:  # noqa: BLE001
    print(json.dumps({"variant": "graph1", "error": repr(ex)[:300]}), flush=True)

for C in (2, 4, 8):
    for prio in (True, False):
        name = f"stag{C}{'p' if prio else ''}"
        try:
            r = time_it(lambda i, C=C, prio=prio: staggered(C, i, prio))
            print(json.dumps({"variant": name, **r}), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"variant": name, "error": repr(ex)[:300]}), flush=True)
        try:
            g, _ = graphed(lambda i, C=C, prio=prio: staggered(C, i, prio))
            r = time_it(lambda i: g.replay())
            print(json.dumps({"variant": "g" + name, **r}), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"variant": "g" + name, "error": repr(ex)[:300]}), flush=True)

# back-to-back throughput without flush (steady-state service): eager two-stream alternation of whole steps
def alt_steps(n, first):
    main = torch.cuda.current_stream()
    st = [s_solve, s_score]
    for s in st:
        s.wait_stream(main)
    for i in range(n):
        with torch.cuda.stream(st[i & 1]):
            engine.ransac_e5_test(matches, logits, K, thr, seed=42, offset=first + i)
    for s in st:
        main.wait_stream(s)


for name, fn in (("serial_40steps", lambda: [eager1(100 + i) for i in range(40)]),
                 ("alt2_40steps", lambda: alt_steps(40, 200))):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"variant": name, "ms_per_step": a.elapsed_time(b) / 40}), flush=True)
