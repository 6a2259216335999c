#!/usr/bin/env python
"""What bounds score_msac_tc_kernel?  Builds csrc/score_tc.cu with each DRB_TC_ABLATE switch set (see the top of that
file) into profiles/microbench/build/libtc_ablate<n>.so and times drb_score_msac_tc on the cfg2 model list.

    python profiles/microbench/tc_ablate.py build        # here (nvcc cross-compiles; the .so files travel with gpurun)
    python profiles/microbench/tc_ablate.py [words...]   # on the B200: one JSON line per (ablation, variant)

The ablated kernels compute garbage; only their duration means anything."""
import ctypes
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "differentiable_ransac_b200", "csrc")
OUT = os.path.join(HERE, "build")
ABLATIONS = {0: "none", 16: "early release of the accumulator", 1: "no MMA", 2: "no epilogue math", 3: "no MMA, no math",
             8: "no MUFU", 32: "4 of 6 K steps", 64: "3 of 6 K steps", 34: "4 of 6 K steps, no epilogue math",
             128: "idle roles poll without the nanosleep back-off",
             1000: "accumulator handed over in halves (DRB_TC_HALF=1)",
             2001: "two-SM kernel: no MMAs issued", 2002: "two-SM kernel: no epilogue arithmetic",
             2003: "two-SM kernel: barriers, copies and tcgen05.ld only"}
if os.environ.get("DRB_ABLATE_ONLY"):
    ABLATIONS = {int(k): ABLATIONS.get(int(k), "?") for k in os.environ["DRB_ABLATE_ONLY"].split(",")}


def build():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for n in ABLATIONS:
        so = os.path.join(OUT, f"libtc_ablate{n}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "--expt-relaxed-constexpr", *([f"-DDRB_TC_ABLATE={n}"] if n < 1000 else ["-DDRB_TC_HALF=1"] if n == 1000 else [f"-DDRB_TCP_ABLATE={n - 2000}"]),
               "-shared", "-o", so, os.path.join(CSRC, "score_tc.cu"), os.path.join(CSRC, "score_tc2.cu"),
               os.path.join(CSRC, "score_tc_pair.cu"), "-lcudart"]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait():
            raise SystemExit("nvcc failed")


def main(variants):
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from differentiable_ransac_b200 import ops
    dev = torch.device("cuda", 0)
    B, K, N = 32, 1000, 2000
    matches_h, logits_h, thr_h, _ = bench.make_inputs(B, N, seed=1234)
    m, lg, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    idx = ops.sample_sets(lg, K, 5, seed=7, offset=0)
    _, _, cm, cid, cc = ops.solve_e5(m, idx, compact=True)
    M = cm.shape[1]
    P = ctypes.c_void_p
    stream = torch.cuda.current_stream().cuda_stream
    ref_best = {}
    for n, what in ABLATIONS.items():
        lib = ctypes.CDLL(os.path.join(OUT, f"libtc_ablate{n}.so"))
        lib.drb_score_msac_tc_workspace_bytes.restype = ctypes.c_size_t
        lib.drb_score_msac_tc.argtypes = [P, P, P, P, P] + [ctypes.c_int] * 4 + [P, P, P, ctypes.c_size_t, P]
        nbytes = int(lib.drb_score_msac_tc_workspace_bytes(B, N))
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
        for name in variants:
            words = ops._TC_WORDS[name]
            best = torch.zeros(B, dtype=torch.int64, device=dev)

            def call():
                rc = lib.drb_score_msac_tc(m.data_ptr(), cm.data_ptr(), cc.data_ptr(), cid.data_ptr(), thr.data_ptr(), B, M, N,
                                           words, None, best.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream)
                assert rc == 0, rc
            for _ in range(3):
                call()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for a, b in ev:
                flush.fill_(1.0)
                a.record()
                call()
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            best.zero_()
            call()
            torch.cuda.synchronize()
            if n == 0:
                ref_best[name] = best.clone()
            same = int((best == ref_best[name]).sum()) if name in ref_best else None
            print(json.dumps(dict(ablate=n, what=what, kernel=name, ms_median=round(ts[len(ts) // 2], 5), ms_min=round(ts[0], 5),
                                  models=int(cc.sum()), same_winners_as_unablated=same)), flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build()
    else:
        main(sys.argv[1:] or ["tc_bf16p", "tc_bf16p_s"])
