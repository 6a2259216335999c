#!/usr/bin/env python
"""Benchmark of the hypothesize-and-score hot path (BASELINE.json metric: hypotheses/sec,
5PC-E, 2k correspondences x 1k hypotheses, 32 pairs per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic pairs: Gumbel top-5 sampling
(in-kernel Philox, fresh offset every step) -> Nister 5-point on every sample -> Sampson/MSAC
score of every model against every correspondence -> arg-max + winner mask.  Prints ONE JSON
line (rank 0).  `value` has the inputs resident in HBM; `e2e` goes through the public API with
pinned-host inputs and a device->host read of the result inside the timed region.

--impl reference times the reference's own algorithm on the host cores: the CPU oracle
(`oracle/`, a line-cited restatement of the reference's PyTorch path, pinned to the reference
by tests/golden) -- /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(B=32, K=1000, N=2000, sample_size=5, slots=10, threshold_px=0.75, focal=800.0)
# compulsory traffic of the whole forward per hypothesis (matches+logits in,
# 10 models + 10 scores out, best index + winner mask) at the headline shape.
ALGO_BYTES_PER_HYP = 442.0
WORKLOAD_NAME = ("cfg2: Essential 5PC (Nister), 32 pairs x 1000 hyps x 2000 corrs per GPU, fwd only, "
                 "test-mode semantics (sample -> solve -> MSAC -> arg-max + winner mask)")
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # non-tensor FP32 FMA peak of a B200 at its 1965 MHz boost: 74.4
REF_BUDGET_S = 150.0      # the reference arm sizes its per-step sample so that the whole run stays below this
ALGO_FLOP_PER_HYP = 0.82e6


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_side(n_models, n_points, kernel_ms, scorer):
    """tcgen05 work of one launch of the tensor-core scorer: every (128 correspondences x 128 models) tile is six
    128 x 256 x 8 (TF32) or x 16 (BF16) MMAs, padded tiles included.  Peak: the driver-measured dense BF16 rate
    (MEASURED_PEAKS.json; TF32 runs at half of it).  Extra keys beside the mandated HBM roofline; never raises."""
    try:
        bf16 = "bf16" in scorer or scorer == "tc"
        tiles = -(-n_points // 128) * -(-n_models // 128)            # lower bound: per-pair tails add a little
        flop = tiles * 6 * 2.0 * 128 * 256 * (16 if bf16 else 8)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["bf16_tflops"])
            src = "measured dense BF16 (MEASURED_PEAKS.json)"
        except Exception:
            peak, src = 2250.0, "nominal dense BF16"
        if not bf16:
            peak, src = peak / 2, src + " / 2 for TF32"
        t = flop / (kernel_ms / 1e3) / 1e12
        return dict(tensor_tflops=t, tensor_peak_tflops=peak, tensor_frac=t / peak, tensor_peak_source=src,
                    tensor_note="flops issued on tcgen05, the 3 (TF32) or 6 (BF16) partial products of the split "
                                "operands included")
    except Exception as e:                                         # reporting only
        return dict(tensor_note=f"unavailable: {e}")


def load_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)[kernel]
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(B, N, seed):
    from differentiable_ransac_b200 import synth

    matches, E_gt, inl = synth.relative_pose_batch(B, N, seed=seed)
    logits = synth.logits_regime(B, N, "L0", seed=seed + 7)
    thr = torch.full((B,), WORKLOAD["threshold_px"] / WORKLOAD["focal"])
    return matches, logits, thr, E_gt


# ---------------------------------------------------------------------------------------------------
def pick_cpu_threads(N):
    """torch's default (= all host cores) is pathological for this path on a many-core box: the
    per-sample 10x10 eigvals loop (nister.py:355-370) and the tiny batched LAPACK calls spend their
    time in thread wake-ups.  Give the CPU side its BEST case: time a short run at several thread
    counts and keep the fastest."""
    from differentiable_ransac_b200 import synth
    from oracle import driver

    cores = os.cpu_count() or 1
    matches, logits, thr, _ = make_inputs(1, N, seed=4321)
    G = synth.gumbel_noise((48, N), seed=2)
    best = (None, float("inf"))
    for nt in sorted({1, 4, 8, 16, 32, cores}):
        if nt > cores:
            continue
        torch.set_num_threads(nt)
        driver.test_loop(matches[0], logits[0], [G[:8]], float(thr[0]))
        t0 = time.perf_counter()
        driver.test_loop(matches[0], logits[0], [G], float(thr[0]))
        dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (nt, dt)
    torch.set_num_threads(best[0])
    return best[0], cores


def cpu_reference_throughput(max_seconds=12.0, max_pairs=8, K=None, N=None):
    """The reference's algorithm on the host cores (oracle port): serial over pairs exactly as
    model_cl.py:488 is, one chunk of K hypotheses per pair (ransac_batch_size = K)."""
    from differentiable_ransac_b200 import synth
    from oracle import driver

    K = K or WORKLOAD["K"]
    N = N or WORKLOAD["N"]
    threads, cores = pick_cpu_threads(N)
    matches, logits, thr, _ = make_inputs(max_pairs, N, seed=1234)
    done, t_total = 0, 0.0
    # warm-up on a small chunk (LAPACK / thread pool initialisation)
    driver.test_loop(matches[0], logits[0], [synth.gumbel_noise((16, N), seed=1)], float(thr[0]))
    for b in range(max_pairs):
        G = synth.gumbel_noise((K, N), seed=100 + b)
        t0 = time.perf_counter()
        driver.test_loop(matches[b], logits[b], [G], float(thr[b]))
        t_total += time.perf_counter() - t0
        done += 1
        if t_total > max_seconds:
            break
    return dict(value=done * K / t_total, unit="hypotheses/s", cores=threads, kind="port",
                sample=f"{done} pair(s) x {K} hyps x {N} corrs, oracle/driver.test_loop (sample+5pt+MSAC), "
                       f"{t_total:.2f} s, torch {torch.__version__} CPU, {threads} threads (fastest of a sweep; "
                       f"host has {cores} cores)")


def auc_parity(dev, pairs=12, N=1000, K=192, scorer=None):
    """AUC@5/10/20 of the poses recovered from the winning E, CUDA path vs the CPU oracle of the reference,
    same synthetic pairs and the same injected Gumbel noise (bounded sample: ~5 s of CPU work)."""
    from differentiable_ransac_b200 import engine, synth
    from oracle import driver, pose_eval

    thr = WORKLOAD["threshold_px"] / WORKLOAD["focal"]
    data = [synth.relative_pose_pair(N, (0.35, 0.5, 0.65)[b % 3], seed=900 + b, noise=4e-4, return_pose=True)
            for b in range(pairs)]
    matches = torch.stack([d[0] for d in data])
    logits = synth.logits_regime(pairs, N, "L0", seed=12)
    noise = synth.gumbel_noise((pairs, K, N), seed=13)
    ours = engine.ransac_e5_test(matches.to(dev), logits.to(dev), K, torch.full((pairs,), thr, device=dev),
                                 noise=noise.to(dev), want_scores=True, scorer=scorer)
    e_ours, e_ref, same = [], [], 0
    for b in range(pairs):
        _, _, _, R, t = data[b]
        ref = driver.test_loop(matches[b], logits[b], [noise[b]], thr)
        same += int(int(ours["best_hyp"][b]) == ref["best_idx"] // 10)
        e_ours.append(max(pose_eval.pose_error_deg(ours["best_model"][b].cpu().numpy(), matches[b].numpy(), R, t,
                                                   ours["mask"][b].cpu().numpy())))
        e_ref.append(max(pose_eval.pose_error_deg(ref["best_model"].numpy(), matches[b].numpy(), R, t,
                                                  ref["best_mask"].numpy())))
    a, r = pose_eval.auc(e_ours), pose_eval.auc(e_ref)
    # the same evaluation without leaving the device (drb_recover_pose: decomposition, DLT cheirality vote over all
    # correspondences as test.py:73-76 does, angular errors), one launch for all pairs
    from differentiable_ransac_b200 import cv_utils
    R_gt = torch.stack([d[3] for d in data]).float().to(dev)
    t_gt = torch.stack([d[4] for d in data]).float().to(dev)
    err = cv_utils.pose_errors(ours["best_model"], matches.to(dev), R_gt, t_gt)[:, 0]
    a_dev = pose_eval.auc(err.max(dim=-1).values.cpu().tolist())
    return dict(auc5_10_20_ours=a, auc5_10_20_cpu_reference=r, auc5_10_20_ours_device_pose=a_dev,
                same_best_hypothesis=f"{same}/{pairs}", scorer=scorer or "default (FP32 work queue)",
                sample=f"{pairs} synthetic pairs x {K} hyps x {N} corrs, identical injected Gumbel noise, "
                       "pose from cv2.recoverPose, AUC as cv_utils.py:528-546")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from differentiable_ransac_b200 import synth
    from oracle import driver

    K, N = WORKLOAD["K"], WORKLOAD["N"]
    cores, host_cores = pick_cpu_threads(N)
    pairs_per_step = 1
    matches, logits, thr, _ = make_inputs(pairs_per_step, N, seed=1234)
    # Bounded sample: a step is 1 pair x Ks hypotheses x N correspondences of the workload (the reference is
    # serial over pairs, model_cl.py:488, and linear in the hypotheses: a Python loop over the samples).  Ks is
    # the workload's 1000 unless --steps is so large that the run would pass REF_BUDGET_S; then it shrinks.
    t0 = time.perf_counter()
    driver.test_loop(matches[0], logits[0], [synth.gumbel_noise((100, N), seed=99)], float(thr[0]))
    per_hyp = (time.perf_counter() - t0) / 100
    total = max(1, args.warmup + args.steps)
    Ks = int(min(K, max(50, REF_BUDGET_S / (total * per_hyp))))
    times = []
    for it in range(args.warmup + args.steps):
        G = synth.gumbel_noise((Ks, N), seed=100 + it)
        t0 = time.perf_counter()
        driver.test_loop(matches[0], logits[0], [G], float(thr[0]))
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = pairs_per_step * Ks / (ms / 1e3)
    sample = (f"each step = {pairs_per_step} pair x {Ks} hyps x {N} corrs of the workload on the host cores "
              f"(oracle/driver.test_loop: sample + 5-point + MSAC + arg-max), {cores} threads (fastest of a sweep; "
              f"host has {host_cores} cores)")
    line = dict(impl="reference", metric="hypotheses_per_sec", value=value, unit="hypotheses/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD_NAME, pairs_per_gpu=WORKLOAD["B"], hypotheses_per_pair=K,
                            correspondences=N, sample=sample),
                cpu_baseline=dict(value=value, unit="hypotheses/s", cores=cores, kind="port",
                                  sample=f"{args.steps} step(s); " + sample),
                e2e=dict(value=value, unit="hypotheses/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    return line


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-slots", type=int, default=int(os.environ.get("DRB_E2E_SLOTS", "3")),
                    help="batches in flight in the end-to-end service (engine.E5TestService)")
    ap.add_argument("--e2e-graph", type=int, default=int(os.environ.get("DRB_E2E_GRAPH", "1")),
                    help="1: each service slot replays one CUDA graph (copy-in, kernels, copy-out) per batch")
    ap.add_argument("--value-graph", type=int, default=int(os.environ.get("DRB_VALUE_GRAPH", "1")),
                    help="1: the device-resident steps replay one CUDA graph per stream")
    ap.add_argument("--value-streams", type=int, default=int(os.environ.get("DRB_VALUE_STREAMS", "3")),
                    help="CUDA streams the K device-resident steps are issued over (1 = strictly serial)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else a library prints there while we run (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference_arm(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def run_ours(args):

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from differentiable_ransac_b200 import engine, ops

    B, K, N = WORKLOAD["B"], WORKLOAD["K"], WORKLOAD["N"]
    warmup = max(args.warmup, 3)
    matches_h, logits_h, thr_h, E_gt = make_inputs(B, N, seed=1234 + 1000 * rank)   # pairs shard over ranks
    matches_h, logits_h, thr_h = matches_h.pin_memory(), logits_h.pin_memory(), thr_h.pin_memory()
    matches, logits, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
    # L2 policy for `value`: every step copies its (packed) inputs from one of NB places in HBM, 168 MB in all
    # (> the 126 MB L2), so a step never finds its inputs cached by an earlier one; the roofline pass below
    # (one kernel timed alone) flushes L2 with a 256 MB write instead.
    NB = 128
    packed_all = torch.cat((matches.flatten(), logits.flatten(), thr)).unsqueeze(0).repeat(NB, 1)   # [NB, n_in]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2
    S = max(1, args.value_streams)
    main_stream = torch.cuda.current_stream()
    dsvc = engine.E5TestService(B, N, K, dev, slots=S, seed=42 + rank, graph=bool(args.value_graph), host_io=False)

    def run_steps(n, first):
        """n independent steps (fresh Philox offset each) through engine.E5TestService(host_io=False): issued
        round-robin over S streams (one CUDA graph per stream when --value-graph 1), so the latency-bound
        5-point kernel of one step overlaps the FMA-bound scoring kernel of the previous one.  Every step
        first copies ITS inputs, device to device, from one of NB distinct places in HBM."""
        fork = torch.cuda.Event()
        fork.record(main_stream)
        dsvc.after(fork)
        for i in range(n):
            dsvc.submit(packed=packed_all[(first + i) % NB])
        dsvc.join(main_stream)
        return dsvc.dev_out[(n - 1) % S]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------------
    out = run_steps(warmup, 0)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(main_stream)
    out = run_steps(args.steps, warmup)
    t_end.record(main_stream)
    barrier()
    ms_local = t_begin.elapsed_time(t_end)
    clock_info = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * K * args.steps / (ms_total / 1e3)

    # the same K steps strictly one after the other on one stream, L2 flushed before each (reported beside it)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        out = engine.ransac_e5_test(matches, logits, K, thr, seed=42 + rank, offset=10_000 + i,
                                    scorer=dsvc.scorer if str(dsvc.scorer).startswith("tc") else None)
        ev[i][1].record()
    barrier()
    ms_serial = sorted(a.elapsed_time(b) for a, b in ev)[args.steps // 2]      # median

    # ---- end to end through the public API with host buffers ("e2e") ------------------------------------
    # engine.E5TestService: every step copies ITS inputs from pinned host memory (one packed buffer: matches |
    # logits | thr) and reads ITS results back (one packed buffer: model | id | score | #inliers) before the
    # slot is reused.  `--e2e-slots` batches are in flight at once, each on its own compute stream, copies on
    # a copy stream: steady-state service throughput.  Inputs come from the host every step, so no L2 flush
    # is inserted here (and it could not be excluded from the timed region once copies and compute overlap).
    svc = engine.E5TestService(B, N, K, dev, slots=args.e2e_slots, seed=42 + rank, graph=bool(args.e2e_graph))
    for s_ in range(svc.slots):
        svc.stage(s_, matches_h, logits_h, thr_h)      # synthetic: the same pairs staged in every slot
    h2d, d2h = svc.h2d_bytes, svc.d2h_bytes

    def run_e2e(n_steps):
        pending = []
        for _ in range(n_steps):
            if len(pending) == svc.slots:
                svc.result(pending.pop(0))              # the caller reads the oldest batch's results (host sync)
            pending.append(svc.submit())
        for s_ in pending:
            svc.result(s_)

    run_e2e(warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    svc.after(e0)
    run_e2e(args.steps)
    svc.join(main_stream)
    e1.record(main_stream)
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K * args.steps / (float(t2.item()) / 1e3)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (score_msac_kernel), timed live with CUDA events -------------------
    idx = ops.sample_sets(logits, K, 5, seed=7, offset=0)
    models, nsol, cm, cid, cc = ops.solve_e5(matches, idx, compact=True)
    n_valid = int(cc.sum().item())
    reps = 10
    msac_kernel = dsvc.scorer or ops._MSAC_KERNEL     # the kernel the timed steps above ran
    msac_name = {"block": "score_msac_kernel", "stream": "score_msac_stream_kernel"}.get(msac_kernel,
                                                                                           "score_msac_tc_kernel")
    launches_per_step = 5 if msac_name == "score_msac_tc_kernel" else 4   # tc: operand images + scorer
    for _ in range(3):
        ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=True, kernel=msac_kernel)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b_ in evs:
        flush.fill_(1.0)
        a.record()
        ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=True, kernel=msac_kernel)
        b_.record()
    torch.cuda.synchronize()
    score_ms = sum(a.elapsed_time(b_) for a, b_ in evs) / reps          # includes the 8-byte/pair memset of `best`
    # stage shares of one step (events around each stage)
    shares = {}
    for name, fn in (("sample_sets", lambda: ops.sample_sets(logits, K, 5, seed=7, offset=1)),
                     ("sample_gumbel_race", lambda: ops.sample(logits, K, 5, 1.0, seed=7, offset=1)),
                     ("solve_e5", lambda: ops.solve_e5(matches, idx, compact=True)),
                     ("score_msac", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=False,
                                                           kernel=msac_kernel)),
                     ("score_msac_stream", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid,
                                                                  want_scores=False, kernel="stream")),
                     ("score_msac_block", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid,
                                                                 want_scores=False, kernel="block")),
                     ("score_msac_tc", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid,
                                                              want_scores=False, kernel="tc")),
                     ):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            fn()
        b_.record()
        torch.cuda.synchronize()
        shares[name + "_ms"] = a.elapsed_time(b_) / 5
    peak, peak_src = load_peaks()
    score_bytes = B * N * 16 + n_valid * (36 + 4 + 4) + B * 8
    achieved = score_bytes / (score_ms / 1e3) / 1e9
    flops_score = n_valid * N * 37.0
    line = dict(
        metric="hypotheses_per_sec", value=value, unit="hypotheses/s", n_gpus=world, steps=args.steps,
        warmup=warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32", data="synthetic",
        config=dict(workload=WORKLOAD_NAME,
                    pairs_per_gpu=B, hypotheses_per_pair=K, correspondences=N,
                    noise="in-kernel Philox4x32-10; sets drawn without replacement from softmax(logits) "
                          "(= Gumbel top-5 in law, drb_sample_sets)",
                    l2="`value`: every step copies its inputs (device to device, inside the timed region) from one "
                       "of 128 distinct places in HBM, 168 MB in all (> L2); "
                       "serial_ms_per_step and the roofline pass: L2 flushed by a 256 MB write before each launch; "
                       "e2e re-copies its inputs from the host every step",
                    value_streams=S, value_graph=bool(args.value_graph), serial_ms_per_step=ms_serial,
                    scorer=f"{msac_name} ({msac_kernel})",
                    e2e_mode=f"engine.E5TestService(graph={bool(args.e2e_graph)}), {args.e2e_slots} batches in flight "
                             "(one stream each; a CUDA graph per slot when graph=True): packed H2D per step, one packed D2H of (model, id, score, #inliers) "
                             "per step, results read on the host before a slot is reused",
                    parallelism=f"pairs sharded over {world} GPU(s)"),
        clocks=clock_info,
        e2e=dict(value=e2e_value, unit="hypotheses/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=launches_per_step * args.steps,
        roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                      traffic=load_traffic(msac_name),
                      kernel=msac_name, kernel_ms=score_ms, algorithmic_bytes=score_bytes,
                      peak_source=peak_src, models_scored=n_valid,
                      note=("FP32-issue bound, not HBM bound (SURVEY H8): see fp32_tflops" if launches_per_step == 4 else
                            "not HBM bound: the contraction runs on tcgen05 (operands split into TF32 / BF16 words), the "
                            "epilogue is SFU-bound (ncu: XU pipe 67 %); fp32_tflops counts the 37 flop per (model, correspondence) of the FP32 formula "
                            "as useful work, so it can exceed the FP32 pipe's peak (DESIGN.md section 10); kernel_ms covers both "
                            "launches of the call (operand images of the correspondences, then the scorer)"),
                      fp32_tflops=flops_score / (score_ms / 1e3) / 1e12,
                      fp32_peak_tflops=FP32_PEAK_TFLOPS,
                      fp32_frac=flops_score / (score_ms / 1e3) / 1e12 / FP32_PEAK_TFLOPS,
                      step_algorithmic_gbs=ALGO_BYTES_PER_HYP * B * K / (ms_per_step / 1e3) / 1e9, **shares),
    )
    if launches_per_step == 5:
        line["roofline"].update(tensor_side(n_valid, N, score_ms, str(msac_kernel)))
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only
        line["cpu_baseline"] = cpu_reference_throughput()
        line["accuracy"] = auc_parity(dev, scorer=dsvc.scorer)      # through the scorer the timed steps used
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
