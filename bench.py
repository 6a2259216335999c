#!/usr/bin/env python
"""Benchmark of the hypothesize-and-score hot path (BASELINE.json metric: hypotheses/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg1|cfg3|cfg4|cfg5]

Default `--config cfg2` is the headline: 5PC-E (Nister), 32 pairs x 1000 hypotheses x 2000 correspondences per
GPU, forward only.  One "step" = one pass of the hot path over one batch of synthetic pairs.  Prints ONE JSON line
(rank 0).  `value` has the inputs resident in HBM; `e2e` goes through the reference-facing plugin
(`model_cl.RANSACLayer`) with pinned-HOST tensors in and host results out, copies inside the timed region.

    cfg1  Essential 5PC Stewenius, 1 pair x 100 hyps x 2000 corrs          (the reference's own CPU-runnable case)
    cfg2  Essential 5PC Nister, 32 x 1000 x 2000, fwd only                 (headline)
    cfg3  Fundamental 8PC, 64 x 2000 x 2000, fwd + bwd training step (epipolar loss)
    cfg4  Rigid 3-point, 16 x 1000 x 50 000, fwd + bwd of the mean residual
    cfg5  Essential 5PC training, 32 pairs per GPU x 1000 x 2000 fwd + bwd, + the NCCL all-reduce of a
          CLNet-sized (2.49 MB) gradient buffer at --gpus > 1

--impl reference times the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py in the build container;
/root/reference does not exist on the GPU box) on the host cores through its own public entry
(`model_cl.RANSACLayer.forward`), falling back to the CPU oracle port when oracle/_ref is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(B=32, K=1000, N=2000, sample_size=5, slots=10, threshold_px=0.75, focal=800.0)
# compulsory traffic of the whole forward per hypothesis (matches+logits in,
# 10 models + 10 scores out, best index + winner mask) at the headline shape.
ALGO_BYTES_PER_HYP = 442.0
WORKLOAD_NAME = ("cfg2: Essential 5PC (Nister), 32 pairs x 1000 hyps x 2000 corrs per GPU, fwd only, "
                 "test-mode semantics (sample -> solve -> MSAC -> arg-max + winner mask)")
CONFIGS = {
    "cfg1": dict(kind="e5", mode="test", B=1, K=100, N=2000,
                 name="cfg1: Essential 5PC Stewenius, 1 pair x 100 hyps x 2000 corrs, loop body (sample -> solve -> MSAC)"),
    "cfg2": dict(kind="e5", mode="test", B=32, K=1000, N=2000, name=WORKLOAD_NAME),
    "cfg3": dict(kind="f8", mode="train", B=64, K=2000, N=2000,
                 name="cfg3: Fundamental 8PC, 64 pairs x 2000 hyps x 2000 corrs per GPU, fwd + bwd training step "
                      "(loss.py epipolar MatchLoss on the GT inliers, gradient to the sampling weights)"),
    "cfg4": dict(kind="rigid", mode="train", B=16, K=1000, N=50000,
                 name="cfg4: Rigid 3-point SVD (3DMatch shape), 16 x 1000 hyps x 50 000 corrs per GPU, train-mode "
                      "forward + backward of the mean squared residual"),
    "cfg5": dict(kind="e5", mode="train", B=32, K=1000, N=2000,
                 name="cfg5: Essential 5PC training, 32 pairs per GPU x 1000 hyps x 2000 corrs (256 pairs over 8 GPUs), "
                      "fwd + bwd (closest-to-GT slot, MatchLoss), NCCL all-reduce of a CLNet-sized gradient buffer"),
}
CLNET_PARAMS = 622_616          # the weight network's parameter count = the all-reduced gradient (2.49 MB)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # non-tensor FP32 FMA peak of a B200 at its 1965 MHz boost: 74.4
XU_PEAK_GOPS = 148 * 16 * 1.965                     # MUFU results per ns: 16 lanes per SM per clock
REF_BUDGET_S = 150.0      # the reference arm sizes its per-step sample so that the whole run stays below this


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
    (MEASURED_PEAKS.json; TF32 runs at half of it).  Also the SFU side: one MUFU.RCP per (model, correspondence), or
    one per two in the pair-reciprocal variants, against 16 results per SM per clock.  Never raises."""
    try:
        bf16 = "bf16" in scorer or scorer == "tc"
        pair = scorer.replace("_e16", "").replace("_s", "").endswith(("p", "q"))
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
        rcp = tiles * 128 * 128 * (0.5 if pair else 1.0)
        xu = rcp / (kernel_ms * 1e6)                                  # results per ns
        return dict(tensor_tflops=t, tensor_peak_tflops=peak, tensor_frac=t / peak, tensor_peak_source=src,
                    tensor_note="flops issued on tcgen05, the 3 (TF32) or 6 (BF16) partial products of the split "
                                "operands included",
                    xu_gops=xu, xu_peak_gops=XU_PEAK_GOPS, xu_frac=xu / XU_PEAK_GOPS,
                    xu_note="MUFU.RCP results per ns against 148 SMs x 16 lanes x 1.965 GHz")
    except Exception as e:                                         # reporting only
        return dict(tensor_note=f"unavailable: {e}")


def load_traffic(kernel, warm=False):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full captures
    (profiles/r2_traffic.json, else round 1's).  Default: ncu's own cache control (L2 flushed before the kernel);
    warm=True: the capture with --cache-control none, when one exists."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)[kernel]
            if warm:
                return int(t["dram_bytes_read_warm_l2"]) + int(t["dram_bytes_write_warm_l2"])
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        except Exception:
            continue
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


class StartGate:
    """Holds the GPU work of a timed region back until the host has enqueued it: `close(stream)` enqueues a stream
    memory operation (cuStreamWaitValue32 on a flag in pinned host memory) that blocks `stream` -- and every stream
    made to wait on it -- until `open()` stores 1 into the flag.  With the steps enqueued behind the gate, the CUDA
    events around them time the device alone: the host's enqueue jitter (a few percent of a 20-step, ~3 ms window,
    and the max over N ranks picks the unluckiest) stays outside.  At most `depth` steps are enqueued before the
    gate opens, far below the driver's launch-queue depth.  Falls back to no gate when the driver call is
    unavailable (DRB_BENCH_GATE=0 turns it off)."""

    depth = 64

    def __init__(self):
        self.ok = os.environ.get("DRB_BENCH_GATE", "1") != "0"
        # under a profiler that serialises kernels (ncu waits for each launch to finish inside the launch call) a
        # closed gate is a deadlock: the host cannot reach open()
        if any(k.startswith(("NV_COMPUTE_PROFILER", "CUDA_INJECTION", "NV_NSIGHT")) for k in os.environ):
            self.ok = False
        try:
            from cuda.bindings import driver as cu

            self.cu = cu
            self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        except Exception:
            self.ok = False

    def close(self, stream):
        if not self.ok:
            return False
        self.flag[0] = 0
        try:
            (err,) = self.cu.cuStreamWaitValue32(stream.cuda_stream, self.flag.data_ptr(), 1,
                                                 self.cu.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_EQ)
            self.closed = err == self.cu.CUresult.CUDA_SUCCESS
        except Exception:
            self.closed = False
        return self.closed

    def open(self):
        if self.ok:
            self.flag[0] = 1


def make_inputs(B, N, seed):
    from differentiable_ransac_b200 import synth

    matches, E_gt, inl = synth.relative_pose_batch(B, N, seed=seed)
    logits = synth.logits_regime(B, N, "L0", seed=seed + 7)
    thr = torch.full((B,), WORKLOAD["threshold_px"] / WORKLOAD["focal"])
    return matches, logits, thr, E_gt


def _inlier_pack(matches, inl):
    B = matches.shape[0]
    npts = inl.sum(1).int()
    P = int(npts.max())
    pts = torch.zeros(B, P, 4)
    for b in range(B):
        pts[b, : int(npts[b])] = matches[b][inl[b]]
    return pts, npts


def make_train_inputs(cfg, B, seed):
    """Synthetic batch of a training config: dict of CPU tensors."""
    from differentiable_ransac_b200 import synth

    kind, N = cfg["kind"], cfg["N"]
    if kind == "e5":
        matches, gt, inl = synth.relative_pose_batch(B, N, seed=seed, noise=2e-4)
        pts, npts = _inlier_pack(matches, inl)
        return dict(matches=matches, logits=synth.logits_regime(B, N, "L0", seed=seed + 4), gt=gt, pts=pts, npts=npts)
    if kind == "f8":
        pairs = [synth.pixel_pair(N, 0.5, seed=seed + b) for b in range(B)]
        matches = torch.stack([p[0] / 640.0 for p in pairs])
        pts, npts = _inlier_pack(matches, torch.stack([p[3] for p in pairs]))
        return dict(matches=matches, logits=synth.logits_regime(B, N, "L1", seed=seed + 1), gt=None, pts=pts, npts=npts,
                    F=torch.stack([p[1] for p in pairs]), pixels=torch.stack([p[0] for p in pairs]))
    pts3 = [synth.rigid_pair(N, 0.7, seed=seed + b) for b in range(B)]
    return dict(matches=torch.stack([p[0] for p in pts3]), logits=synth.logits_regime(B, N, "L1", seed=seed + 2), gt=None,
                pts=None, npts=None, pose=torch.stack([p[1] for p in pts3]))


# ---------------------------------------------------------------------------------------------------
# the CPU side: the unmodified reference (oracle/_ref) when present, else the oracle port
def _harness():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness

    return ref_harness if ref_harness.reference_dir() is not None else None


def pick_cpu_threads_port(N):
    """Best thread count for the oracle PORT (used only when oracle/_ref is absent)."""
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


def _cpu_step_fn(cfg_name, Ks, pair=0, seed=1234):
    """-> (callable running ONE pair x Ks hypotheses of the config's workload on the host cores, kind, entry)."""
    cfg = CONFIGS[cfg_name]
    rh = _harness()
    N = cfg["N"]
    if cfg["mode"] == "test":
        matches, logits, thr, _ = make_inputs(pair + 1, N, seed=seed)
        m, w, th = matches[pair], logits[pair], float(thr[pair])
        if rh is not None:
            if cfg_name == "cfg1":
                return (lambda: rh.stewenius_loop_body(m, w, Ks, th)), "reference", \
                    "ransac.py:55-144 loop body with EssentialMatrixEstimator (Stewenius), unmodified reference"
            K1, K2, im1, im2 = rh.intrinsics()
            layer = rh.make_layer(Ks)

            def run():
                with torch.no_grad():
                    layer.forward(m, w, K1, K2, im1, im2)
            return run, "reference", "model_cl.RANSACLayer.forward, unmodified reference"
        from differentiable_ransac_b200 import synth
        from oracle import driver

        solver = "stewenius" if cfg_name == "cfg1" else "nister"
        return (lambda: driver.test_loop(m, w, [synth.gumbel_noise((Ks, N), seed=100)], th, solver=solver)), "port", \
            "oracle/driver.test_loop (sample + 5-point + MSAC + arg-max)"
    data = make_train_inputs(cfg, pair + 1, seed)
    if rh is None:
        raise RuntimeError("the training configs need oracle/_ref (python oracle/make_ref.py in the build container)")
    kind = cfg["kind"]
    if kind == "f8":
        pix = data["pixels"][pair]
        pts = pix.clone()
        c = torch.tensor([320.0, 240.0])
        pts[:, 0:2] = (pix[:, 0:2] - c) / 640.0          # datasets.py:74-79: (pixels - centre) / max(im)
        pts[:, 2:4] = (pix[:, 2:4] - c) / 640.0
        gt = data["F"][pair]
    else:
        pts = data["matches"][pair]
        gt = data["gt"][pair] if kind == "e5" else data["pose"][pair]
    w = data["logits"][pair]
    return (lambda: rh.train_step(kind, Ks, pts, w, gt)), "reference", \
        {"e5": "RANSACLayer(train).forward + MatchLoss + backward", "f8": "RANSACLayer(-fmat 1 -sam 3, train).forward + "
         "MatchLoss + backward", "rigid": "RANSACLayer3D(train).forward + loss.backward"}[kind] + ", unmodified reference"


def _cpu_threads(cfg_name):
    rh = _harness()
    N = min(CONFIGS[cfg_name]["N"], 2000)
    if rh is not None:
        matches, logits, _, _ = make_inputs(1, N, seed=4321)
        return rh.pick_threads(matches, logits)
    return pick_cpu_threads_port(N)


def cpu_reference_throughput(cfg_name="cfg2", max_seconds=12.0):
    """The reference on the host cores, a BOUNDED sample of the config's workload: pairs are processed one after the
    other exactly as model_cl.py:488 does, one chunk of K hypotheses per pair (fewer when one pair would not fit
    the time budget: the reference's loop is linear in the hypotheses)."""
    cfg = CONFIGS[cfg_name]
    threads, cores = _cpu_threads(cfg_name)
    K, N = cfg["K"], cfg["N"]
    probe_K = max(8, min(K, 32))
    fn, kind, entry = _cpu_step_fn(cfg_name, probe_K)
    fn()                                                     # warm-up (imports, LAPACK, thread pool)
    t0 = time.perf_counter()
    fn()
    per_hyp = (time.perf_counter() - t0) / probe_K
    Ks = int(max(8, min(K, max_seconds / max(per_hyp, 1e-9))))
    done, total, pair = 0, 0.0, 0
    while total < max_seconds and pair < 8:
        fn, kind, entry = _cpu_step_fn(cfg_name, Ks, pair=pair)
        t0 = time.perf_counter()
        fn()
        total += time.perf_counter() - t0
        done += Ks
        pair += 1
    out = dict(value=done / total, unit="hypotheses/s", cores=threads, kind=kind,
               sample=f"{pair} pair(s) x {Ks} hyps x {N} corrs of {cfg_name}, {entry}, {total:.2f} s, torch "
                      f"{torch.__version__} CPU, {threads} threads (fastest of a sweep; host has {cores} cores)")
    rh = _harness()
    if rh is not None:
        out["reference"] = rh.provenance()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    K, N = cfg["K"], cfg["N"]
    cores, host_cores = _cpu_threads(args.config)
    # Bounded sample: a step is 1 pair x Ks hypotheses x N correspondences of the workload (the reference is
    # serial over pairs, model_cl.py:488, and linear in the hypotheses: a Python loop over the samples).  Ks is
    # the workload's K unless --steps is so large that the run would pass REF_BUDGET_S; then it shrinks.
    probe_K = max(8, min(K, 32))
    fn, kind, entry = _cpu_step_fn(args.config, probe_K)
    fn()
    t0 = time.perf_counter()
    fn()
    per_hyp = (time.perf_counter() - t0) / probe_K
    total = max(1, args.warmup + args.steps)
    Ks = int(min(K, max(min(K, 16), REF_BUDGET_S / (total * per_hyp))))
    fn, kind, entry = _cpu_step_fn(args.config, Ks)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = Ks / (ms / 1e3)
    sample = (f"each step = 1 pair x {Ks} hyps x {N} corrs of the workload on the host cores ({entry}), {cores} "
              f"threads (fastest of a sweep; host has {host_cores} cores)")
    rh = _harness()
    line = dict(impl="reference", metric="hypotheses_per_sec", value=value, unit="hypotheses/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=cfg["name"], pairs_per_gpu=cfg["B"], hypotheses_per_pair=K,
                            correspondences=N, sample=sample),
                cpu_baseline=dict(value=value, unit="hypotheses/s", cores=cores, kind=kind,
                                  sample=f"{args.steps} step(s); " + sample),
                e2e=dict(value=value, unit="hypotheses/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    if rh is not None:
        line["cpu_baseline"]["reference"] = rh.provenance()
    return line


def auc_parity(dev, pairs=12, N=1000, K=192, scorer=None):
    """AUC@5/10/20 of the poses recovered from the winning E, CUDA path vs the CPU oracle of the reference,
    same synthetic pairs and the same injected Gumbel noise (bounded sample: ~5 s of CPU work).  Every pair whose
    winning hypothesis differs from the oracle's is classified: a tie (scores within 1e-4 relative) or worse."""
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
    e_ours, e_ref, same, ties, worse, worst = [], [], 0, 0, 0, 0.0
    for b in range(pairs):
        _, _, _, R, t = data[b]
        ref = driver.test_loop(matches[b].double(), logits[b].double(), [noise[b].double()], thr)   # fp64 oracle
        rel = abs(float(ours["best_score"][b]) - float(ref["best_score"])) / float(ref["best_score"])
        if int(ours["best_hyp"][b]) == ref["best_idx"] // 10:
            same += 1
        elif rel <= 1e-4:
            ties += 1
        else:
            worse += 1
        worst = max(worst, rel)
        e_ours.append(max(pose_eval.pose_error_deg(ours["best_model"][b].cpu().numpy(), matches[b].numpy(), R, t,
                                                   ours["mask"][b].cpu().numpy())))
        e_ref.append(max(pose_eval.pose_error_deg(ref["best_model"].float().numpy(), matches[b].numpy(), R, t,
                                                  ref["best_mask"].numpy())))
    a, r = pose_eval.auc(e_ours), pose_eval.auc(e_ref)
    from differentiable_ransac_b200 import cv_utils
    R_gt = torch.stack([d[3] for d in data]).float().to(dev)
    t_gt = torch.stack([d[4] for d in data]).float().to(dev)
    err = cv_utils.pose_errors(ours["best_model"], matches.to(dev), R_gt, t_gt)[:, 0]
    a_dev = pose_eval.auc(err.max(dim=-1).values.cpu().tolist())
    return dict(auc5_10_20_ours=a, auc5_10_20_cpu_reference=r, auc5_10_20_ours_device_pose=a_dev,
                same_best_hypothesis=f"{same}/{pairs}", ties_within_1e_4=ties, worse_than_1e_4=worse,
                worst_rel_score_gap=worst, scorer=scorer or "default (FP32 work queue)",
                sample=f"{pairs} synthetic pairs x {K} hyps x {N} corrs, identical injected Gumbel noise, fp64 CPU oracle "
                       "of the reference loop, pose from cv2.recoverPose, AUC as cv_utils.py:528-546; the full-size "
                       "(32 x 1000 x 2000) index parity of the timed pipeline is tests/test_gpu_pipeline_parity.py")


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-slots", type=int, default=int(os.environ.get("DRB_E2E_SLOTS", "4")),
                    help="batches in flight in the end-to-end plugin calls (RANSACLayer.submit / collect)")
    ap.add_argument("--e2e-graph", type=int, default=int(os.environ.get("DRB_E2E_GRAPH", "1")),
                    help="1: each service slot replays one CUDA graph (kernels, copy-out) per batch")
    ap.add_argument("--value-graph", type=int, default=int(os.environ.get("DRB_VALUE_GRAPH", "1")),
                    help="1: the device-resident steps replay one CUDA graph per stream")
    ap.add_argument("--value-streams", type=int, default=int(os.environ.get("DRB_VALUE_STREAMS", "4")),
                    help="CUDA streams the K device-resident steps are issued over (1 = strictly serial)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else a library prints there while we run (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            line = run_reference_arm(args)
        elif CONFIGS[args.config]["mode"] == "train":
            line = run_train(args)
        else:
            line = run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _setup_dist():
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
    return rank, world, local_rank, dev, dist


def _barrier(dist):
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(ms, dev, dist):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _time_stage(fn, reps=5):
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b_.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b_) / reps


def run_ours(args):
    rank, world, local_rank, dev, dist = _setup_dist()
    from differentiable_ransac_b200 import engine, ops
    from differentiable_ransac_b200.model_cl import RANSACLayer

    cfg = CONFIGS[args.config]
    B, K, N = cfg["B"], cfg["K"], cfg["N"]
    warmup = max(args.warmup, 3)
    # Weak scaling: every rank gets the SAME 32 synthetic pairs and its own Philox stream (seed 42 + rank), so that all
    # ranks do statistically identical work.  (Round 1 seeded the pairs per rank: the number of real five-point
    # solutions, hence the scorer's work, then differs by ~2 % between ranks, and the max over ranks reads as a
    # scaling loss that is really a workload difference.)
    matches_h, logits_h, thr_h, E_gt = make_inputs(B, N, seed=1234)
    matches_h, logits_h, thr_h = matches_h.pin_memory(), logits_h.pin_memory(), thr_h.pin_memory()
    matches, logits, thr = matches_h.to(dev), logits_h.to(dev), thr_h.to(dev)
    # L2 policy for `value`: every step copies its (packed) inputs from one of NB places in HBM, 168 MB in all
    # (> the 126 MB L2), so a step never finds its inputs cached by an earlier one; the roofline pass below
    # (one kernel timed alone) flushes L2 with a 256 MB write instead.
    n_in = B * N * 5 + B
    NB = max(4, -(-168 * 1024 * 1024 // (n_in * 4)))
    NB = min(NB, 4096)
    packed_all = torch.cat((matches.flatten(), logits.flatten(), thr)).unsqueeze(0).repeat(NB, 1)   # [NB, n_in]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2
    S = max(1, args.value_streams)
    main_stream = torch.cuda.current_stream()
    dsvc = engine.E5TestService(B, N, K, dev, slots=S, seed=42 + rank, graph=bool(args.value_graph), host_io=False)

    gate = StartGate()

    def run_steps(n, first, gated=False):
        """n independent steps (fresh Philox offset each) through engine.E5TestService(host_io=False): issued
        round-robin over S streams (one CUDA graph per stream when --value-graph 1), so the latency-bound
        5-point kernel of one step overlaps the scoring kernel of the previous one.  Every step first copies ITS
        inputs, device to device, from one of NB distinct places in HBM.  `gated`: the first steps (up to
        StartGate.depth) are enqueued behind the start gate, which the caller has closed on the main stream."""
        for i in range(n):
            if gated and i == StartGate.depth:
                gate.open()
            dsvc.submit(packed=packed_all[(first + i) % NB])
        dsvc.join(main_stream)
        return dsvc.dev_out[(n - 1) % S]

    # ---- device-resident throughput ("value") -------------------------------------------------------
    run_steps(warmup, 0)
    _barrier(dist)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(dist)
    gated = gate.close(main_stream)          # the slots' streams wait on the main stream (E5TestService.submit)
    t_begin.record(main_stream)
    run_steps(args.steps, warmup, gated=gated)
    t_end.record(main_stream)
    gate.open()
    _barrier(dist)
    ms_total = _max_over_ranks(t_begin.elapsed_time(t_end), dev, dist)
    clock_info = clocks.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * B * K * args.steps / (ms_total / 1e3)

    # the same K steps strictly one after the other on one stream, L2 flushed before each (reported beside it)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(float(i))
        ev[i][0].record()
        engine.ransac_e5_test(matches, logits, K, thr, seed=42 + rank, offset=10_000 + i,
                              scorer=dsvc.scorer if str(dsvc.scorer).startswith("tc") else None)
        ev[i][1].record()
    _barrier(dist)
    ms_serial = sorted(a.elapsed_time(b) for a, b in ev)[args.steps // 2]      # median

    # ---- end to end through the reference-facing plugin with host buffers ("e2e") ------------------------
    # model_cl.RANSACLayer configured for the hot path proper (test mode, every hypothesis scored: opt.adaptive =
    # False, opt.final_refit = False, one chunk of K hypotheses): `submit(points, weights, K1, K2)` takes the
    # batch's pinned HOST tensors, `collect()` returns list_B[E], the winners' inlier masks [B,N] and scores on
    # the host.  `--e2e-slots` batches are in flight, so the copies of one batch overlap the kernels of another;
    # every step's H2D (matches, weights, thresholds) and D2H (models, ids, scores, #inliers, masks) are inside the
    # timed region.  Inputs come from the host every step, so no L2 flush is inserted here.
    opt = types.SimpleNamespace(device=dev, fmat=0, sampler=2, precision=1, tr=0, threshold=WORKLOAD["threshold_px"],
                                ransac_batch_size=K, weighted=0, adaptive=False, final_refit=False, seed=42 + rank)
    layer = RANSACLayer(opt)
    layer.estimator.max_iterations = K
    Kmat = torch.tensor([[WORKLOAD["focal"], 0.0, 320.0], [0.0, WORKLOAD["focal"], 240.0], [0.0, 0.0, 1.0]])
    K1_h = Kmat.expand(B, 3, 3).contiguous()
    slots = max(1, args.e2e_slots)

    def run_e2e(n_steps):
        pending, last = [], None
        for _ in range(n_steps):
            if len(pending) == slots:
                last = layer.collect(pending.pop(0))     # the caller reads the oldest batch's results (host sync)
            pending.append(layer.submit(matches_h, logits_h, K1_h, K1_h, slots=slots, graph=bool(args.e2e_graph)))
        for s_ in pending:
            last = layer.collect(s_)
        return last

    run_e2e(warmup)
    svc = layer._svc
    h2d, d2h = svc.h2d_bytes, svc.d2h_bytes
    _barrier(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record(main_stream)
    svc.after(e0)
    Es, masks, scores = run_e2e(args.steps)
    svc.join(main_stream)
    e1.record(main_stream)
    torch.cuda.synchronize()
    t_host = (time.perf_counter() - t_host0) * 1e3
    _barrier(dist)
    e2e_ms = _max_over_ranks(max(e0.elapsed_time(e1), t_host), dev, dist)
    e2e_value = world * B * K * args.steps / (e2e_ms / 1e3)
    # e2e_check: the plugin call returns what a direct engine call returns for the same Philox position
    chk = RANSACLayer(types.SimpleNamespace(**{**vars(opt)}))
    chk.estimator.max_iterations = K
    Es0, masks0, scores0 = chk.collect(chk.submit(matches_h, logits_h, K1_h, K1_h, slots=1, graph=False))
    direct = engine.ransac_e5_test(matches, logits, K, thr, seed=42 + rank, offset=0, scorer=chk._svc.scorer)
    torch.cuda.synchronize()
    e2e_check = dict(models_equal=bool(torch.equal(torch.stack(Es0), direct["best_model"].cpu())),
                     masks_equal=bool(torch.equal(masks0, direct["mask"].cpu())),
                     scores_equal=bool(torch.equal(scores0, direct["best_score"].cpu())),
                     inliers_in_last_batch=int(masks.sum()))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the MSAC scorer), timed live with CUDA events -------------------
    idx = ops.sample_sets(logits, K, 5, seed=7, offset=0)
    models, nsol, cm, cid, cc = ops.solve_e5(matches, idx, compact=True)
    n_valid = int(cc.sum().item())
    reps = 10
    msac_kernel = dsvc.scorer or ops._MSAC_KERNEL     # the kernel the timed steps above ran
    msac_name = {"block": "score_msac_kernel", "stream": "score_msac_stream_kernel"}.get(msac_kernel,
                                                                                           "score_msac_tc_kernel")
    is_tc = msac_name == "score_msac_tc_kernel"
    launches_per_step = 5 if is_tc else 4   # tc: operand images + scorer
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
    shares = {}
    for name, fn in (("sample_sets", lambda: ops.sample_sets(logits, K, 5, seed=7, offset=1)),
                     ("solve_e5", lambda: ops.solve_e5(matches, idx, compact=True)),
                     ("score_msac", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=False,
                                                           kernel=msac_kernel)),
                     ("score_msac_stream", lambda: ops.score_msac(matches, cm, thr, count=cc, ids=cid,
                                                                  want_scores=False, kernel="stream")),
                     ):
        shares[name + "_ms"] = _time_stage(fn)
    peak, peak_src = load_peaks()
    score_bytes = B * N * 16 + n_valid * (36 + 4 + 4) + B * 8
    achieved = score_bytes / (score_ms / 1e3) / 1e9
    line = dict(
        metric="hypotheses_per_sec", value=value, unit="hypotheses/s", n_gpus=world, steps=args.steps,
        warmup=warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32", data="synthetic",
        config=dict(workload=cfg["name"],
                    pairs_per_gpu=B, hypotheses_per_pair=K, correspondences=N,
                    noise="in-kernel Philox4x32-10; sets drawn without replacement from softmax(logits) "
                          "(= Gumbel top-5 in law, drb_sample_sets)",
                    l2=f"`value`: every step copies its inputs (device to device, inside the timed region) from one "
                       f"of {NB} distinct places in HBM, {NB * n_in * 4 >> 20} MB in all (> L2); "
                       "serial_ms_per_step and the roofline pass: L2 flushed by a 256 MB write before each launch; "
                       "e2e re-copies its inputs from the host every step",
                    value_streams=S, value_graph=bool(args.value_graph), serial_ms_per_step=ms_serial,
                    start_gate=("the timed steps (the first %d of them) are enqueued behind a cuStreamWaitValue32 gate "
                                "that opens when the host has enqueued them: the events time the device, not the host's "
                                "enqueue jitter" % min(args.steps, StartGate.depth)) if gated else "off",
                    scorer=f"{msac_name} ({msac_kernel})",
                    e2e_mode=f"model_cl.RANSACLayer.submit / collect (the reference-facing plugin; test mode, adaptive "
                             f"exit and final refit off = the hot path proper), {slots} batches in flight (one stream "
                             f"each; graph={bool(args.e2e_graph)}): H2D of the caller's pinned matches / weights / "
                             "thresholds, D2H of (model, id, score, #inliers) + the winner's inlier mask [B,N] per step, "
                             "results read on the host before a slot is reused; timed as max(device events, host clock)",
                    e2e_check=e2e_check,
                    parallelism=f"pairs sharded over {world} GPU(s), no collective in the forward; every rank runs the same "
                                "32 synthetic pairs with its own Philox stream (identical work per rank)"),
        clocks=clock_info,
        e2e=dict(value=e2e_value, unit="hypotheses/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=launches_per_step * args.steps,
        roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                      traffic=load_traffic(msac_name if not is_tc else f"score_msac_tc_kernel<{msac_kernel}>"),
                      traffic_warm_l2=load_traffic(f"score_msac_tc_kernel<{msac_kernel}>", warm=True) if is_tc else None,
                      traffic_note="`traffic` is ncu's cold-L2 figure (it flushes the L2 before the kernel: the operand "
                                   "images and models the previous launches of the step left in L2 come back from HBM); "
                                   "`traffic_warm_l2` is the same launch captured with --cache-control none",
                      kernel=msac_name, kernel_ms=score_ms, algorithmic_bytes=score_bytes,
                      peak_source=peak_src, models_scored=n_valid,
                      note=("FP32-issue bound, not HBM bound (SURVEY H8)" if not is_tc else
                            "not HBM bound (SURVEY H8: ~1 900 flop per algorithmic byte): the contraction runs on tcgen05 "
                            "(operands split into BF16 / TF32 words), the epilogue on the SFU and FMA pipes; the "
                            "informative fractions are tensor_frac and xu_frac; kernel_ms covers both launches of the "
                            "call (operand images of the correspondences, then the scorer)"),
                      step_algorithmic_gbs=ALGO_BYTES_PER_HYP * B * K / (ms_per_step / 1e3) / 1e9, **shares),
    )
    if is_tc:
        line["roofline"].update(tensor_side(n_valid, N, score_ms, str(msac_kernel)))
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only
        line["cpu_baseline"] = cpu_reference_throughput(args.config)
        line["accuracy"] = auc_parity(dev, scorer=dsvc.scorer)      # through the scorer the timed steps used
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


# ---------------------------------------------------------------------------------------------------
def _train_stage_table(cfg, data, dev):
    """Per-kernel times of one training step (CUDA events, each stage alone) and the algorithmic bytes of each:
    -> (stages dict name -> (ms, bytes), dominant name)."""
    from differentiable_ransac_b200 import ops

    kind, B, K, N = cfg["kind"], cfg["B"], cfg["K"], cfg["N"]
    s = {"e5": 5, "f8": 8, "rigid": 3}[kind]
    D = 6 if kind == "rigid" else 4
    m, lg = data["matches"].to(dev), data["logits"].to(dev)
    idx, lse, sel_key, _ = ops.sample(lg, K, s, 1.0, None, 3, 0, want_lse=True)
    st = {}
    st["sample"] = (_time_stage(lambda: ops.sample(lg, K, s, 1.0, None, 3, 0, want_lse=True)),
                    B * N * 4 + B * K * (4 * s + 4 + 4 * s))
    if kind == "e5":
        gt = data["gt"].to(dev)
        sel, used, _, _ = ops.solve_e5_select(m, idx, gt)
        valid = sel >= 0
        st["solve_e5_select"] = (_time_stage(lambda: ops.solve_e5_select(m, idx, gt)), B * K * (s * 16 + 36 + 4 + 4))
    elif kind == "f8":
        used, valid = ops.solve_f8(m, idx)
        valid = valid.bool()
        st["solve_f8"] = (_time_stage(lambda: ops.solve_f8(m, idx)), B * K * (s * 16 + 36 + 1))
    else:
        used, valid = ops.solve_rigid3(m, idx, True)
        valid = valid.bool()
        used = torch.where(valid[..., None, None], used, torch.zeros_like(used))
        st["solve_rigid3"] = (_time_stage(lambda: ops.solve_rigid3(m, idx, True)), B * K * (s * 24 + 64 + 1))
    g_row = torch.full((B, K), 1.0 / (B * K), device=dev)
    if kind == "rigid":
        st["rigid_residual_fwd_bwd"] = (_time_stage(lambda: ops.rigid_residual_forward_backward(m, used, g_row)),
                                        B * N * 24 + B * K * (64 + 4 + 4 + 64))
        _, g_used = ops.rigid_residual_forward_backward(m, used, g_row)
        st["solve_rigid3_bwd"] = (_time_stage(lambda: ops.solve_rigid3_backward(m, idx, g_used.reshape(B, K, 16), True)),
                                  B * K * (s * 24 + 64 + s * 24))
        g_pts = ops.solve_rigid3_backward(m, idx, g_used.reshape(B, K, 16), True)
    else:
        pts, npts = data["pts"].to(dev), data["npts"].to(dev)
        P = pts.shape[1]
        st["episym_fwd_bwd"] = (_time_stage(lambda: ops.episym_forward_backward(pts, used, g_row, npts, valid)),
                                B * P * 16 + B * K * (36 + 4 + 4 + 36))
        _, g_used = ops.episym_forward_backward(pts, used, g_row, npts, valid)
        if kind == "e5":
            st["solve_e5_bwd"] = (_time_stage(lambda: ops.solve_e5_backward_chosen(m, idx, used, sel, g_used.reshape(B, K, 9))),
                                  B * K * (s * 16 + 36 + 36 + s * 16))
            g_pts = ops.solve_e5_backward_chosen(m, idx, used, sel, g_used.reshape(B, K, 9))
        else:
            st["solve_f8_bwd"] = (_time_stage(lambda: ops.solve_f8_backward(m, idx, g_used.reshape(B, K, 9), used)),
                                  B * K * (s * 16 + 36 + 36 + s * 16))
            g_pts = ops.solve_f8_backward(m, idx, g_used.reshape(B, K, 9), used)
    st["gather_bwd"] = (_time_stage(lambda: ops.gather_backward(m, idx, g_pts, want_grad_matches=False)),
                        B * K * s * (4 * D + 4 + 4))
    g_sel, _ = ops.gather_backward(m, idx, g_pts, want_grad_matches=False)
    st["sample_bwd"] = (_time_stage(lambda: ops.sample_backward(lg, idx, lse, sel_key, g_sel, 1.0, None, 3, 0)),
                        B * N * 8 + B * K * (4 + 12 * s))
    dominant = max(st, key=lambda k_: st[k_][0])
    return st, dominant


def run_train(args):
    """cfg3 / cfg4 / cfg5: one step = forward + loss + backward for one batch through engine.TrainStep (the fused,
    CUDA-graph form of the autograd path; tests/test_gpu_train_step.py pins the two to each other)."""
    rank, world, local_rank, dev, dist = _setup_dist()
    from differentiable_ransac_b200 import engine

    cfg = CONFIGS[args.config]
    kind, B, K, N = cfg["kind"], cfg["B"], cfg["K"], cfg["N"]
    warmup = max(args.warmup, 3)
    data = make_train_inputs(cfg, B, seed=300)      # the same batch on every rank, its own Philox stream (see run_ours)
    P = None if data["pts"] is None else data["pts"].shape[1]
    step = engine.TrainStep(kind, B, N, K, dev, P=P, seed=5 + rank, graph=True)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
    devd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    in_bytes = sum(devd[k].numel() * devd[k].element_size() for k in ("matches", "logits"))
    # L2 policy: the step's inputs rotate over NB distinct copies in HBM, > 126 MB in all
    NB = max(2, min(256, -(-160 * 1024 * 1024 // in_bytes)))
    rot_m = devd["matches"].unsqueeze(0).repeat(NB, *([1] * devd["matches"].dim()))
    rot_l = devd["logits"].unsqueeze(0).repeat(NB, 1, 1)
    step.load(devd["matches"], devd["logits"], devd.get("gt"), devd["pts"], devd["npts"])
    # cfg5 at --gpus > 1: the data-parallel all-reduce of the weight network's gradients (d loss /
    # d logits is consumed locally by CLNet's backward; only the resulting 622 616 parameter gradients are shared).
    # CLNet is outside this path, so the buffer is a stand-in of its size; the collective runs on its own stream
    # behind the step's backward and the NEXT step waits for it (it needs the updated weights), i.e. it is on the
    # critical path exactly as in train.py -- not hidden behind the next step.
    do_ar = args.config == "cfg5" and dist is not None
    grads = torch.zeros(CLNET_PARAMS, device=dev)
    comm = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    def one_step(i, rotate=True):
        if rotate:
            step.matches.copy_(rot_m[i % NB], non_blocking=True)
            step.logits.copy_(rot_l[i % NB], non_blocking=True)
        loss, gl = step.run()
        if do_ar:
            comm.wait_stream(main_stream)
            with torch.cuda.stream(comm):
                dist.all_reduce(grads)
            main_stream.wait_stream(comm)
        return loss

    for i in range(warmup):
        one_step(i)
    _barrier(dist)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate = StartGate()
    _barrier(dist)
    gated = gate.close(main_stream)
    t0.record()
    for i in range(args.steps):
        if gated and i == StartGate.depth:
            gate.open()
        one_step(warmup + i)
    t1.record()
    gate.open()
    _barrier(dist)
    ms_total = _max_over_ranks(t0.elapsed_time(t1), dev, dist)
    clock_info = clocks.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * B * K * args.steps / (ms_total / 1e3)

    ar_ms = None
    if do_ar:                                         # the collective alone, device-timed, max over ranks
        for _ in range(5):
            dist.all_reduce(grads)
        _barrier(dist)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(50):
            dist.all_reduce(grads)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = _max_over_ranks(a0.elapsed_time(a1) / 50, dev, dist)

    # ---- e2e: pinned host batch in, loss on the host, every step ------------------------------------
    # A training loop is sequential in its steps (step i + 1 needs the weights step i produced), but not in its DATA: the
    # next batch is known.  As a DataLoader with pinned memory does, the batch of step i + 1 is copied host -> device on
    # a copy stream into one of two staging sets while step i computes; the step itself starts with a device-to-device
    # copy of its staged batch into the graph's static buffers, and the host reads the step's loss before it goes on.
    loss_h = torch.empty(1).pin_memory()
    keys = [k_ for k_ in ("matches", "logits", "gt", "pts", "npts") if torch.is_tensor(host.get(k_))]
    stage = [{k_: torch.empty_like(devd[k_]) for k_ in keys} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for e_ in ev_free:
        e_.record(main_stream)

    def prefetch(j):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[j % 2])          # the step that used this staging set has copied it out
            for k_ in keys:
                stage[j % 2][k_].copy_(host[k_], non_blocking=True)
            ev_ready[j % 2].record(copy_stream)

    def e2e_step(j):
        prefetch(j + 1)
        main_stream.wait_event(ev_ready[j % 2])
        st_ = stage[j % 2]
        step.load(st_["matches"], st_["logits"], st_.get("gt"), st_.get("pts"), st_.get("npts"))
        ev_free[j % 2].record(main_stream)
        loss = one_step(0, rotate=False)
        loss_h.copy_(loss, non_blocking=True)
        main_stream.synchronize()                           # the training loop reads the loss (train.py:175)
        return float(loss_h)

    prefetch(0)
    for j in range(warmup):
        e2e_step(j)
    _barrier(dist)
    th0 = time.perf_counter()
    for j in range(warmup, warmup + args.steps):
        last_loss = e2e_step(j)
    torch.cuda.synchronize()
    e2e_ms = _max_over_ranks((time.perf_counter() - th0) * 1e3, dev, dist)
    e2e_value = world * B * K * args.steps / (e2e_ms / 1e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("matches", "logits", "gt", "pts", "npts")
              if torch.is_tensor(host.get(k)))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    stages, dominant = _train_stage_table(cfg, data, dev)
    peak, peak_src = load_peaks()
    dom_ms, dom_bytes = stages[dominant]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    D = 6 if kind == "rigid" else 4
    step_bytes = B * N * (4 * D + 4) + B * K * (64 if kind == "rigid" else 36) + B * N * 4
    line = dict(
        metric="hypotheses_per_sec", value=value, unit="hypotheses/s", n_gpus=world, steps=args.steps, warmup=warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=cfg["name"], pairs_per_gpu=B, hypotheses_per_pair=K, correspondences=N,
                    step="engine.TrainStep: sample (Gumbel top-s, in-kernel Philox, device-side stream position) -> minimal "
                         "solver -> loss -> IFT adjoint of the solver -> gather backward -> straight-through sampler "
                         "backward; ONE CUDA graph replay per step; gradient: d loss / d logits [B,N]",
                    l2=f"the step's inputs rotate over {NB} distinct copies in HBM ({NB * in_bytes >> 20} MB > L2), copied "
                       "device to device inside the timed region",
                    start_gate="on (see cfg2)" if gated else "off",
                    allreduce=(None if not do_ar else
                               dict(bytes=CLNET_PARAMS * 4, ms_alone=ar_ms, share_of_step=ar_ms / ms_per_step,
                                    placement="own stream, behind the step's backward; the next step waits for it "
                                              "(critical path, as in data-parallel train.py)")),
                    e2e_mode="every step: H2D of the NEXT step's pinned host batch on a copy stream into one of two staging sets "
                             "(a DataLoader's prefetch) while this step runs; the step = device-to-device copy of its staged "
                             "batch into the static buffers, graph replay, D2H of "
                             "the loss, host synchronisation (a training loop is sequential: no batches in flight); timed "
                             "on the host clock",
                    last_loss=last_loss, parallelism=f"pairs sharded over {world} GPU(s)"),
        clocks=clock_info,
        e2e=dict(value=e2e_value, unit="hypotheses/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4),
        gpu_launches=len(stages) * args.steps,
        roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                      kernel=dominant, kernel_ms=dom_ms, algorithmic_bytes=dom_bytes, peak_source=peak_src,
                      stages_ms={k_: v[0] for k_, v in stages.items()},
                      step_algorithmic_bytes_per_hyp=step_bytes / (B * K),
                      step_algorithmic_gbs=step_bytes / (ms_per_step / 1e3) / 1e9,
                      note="latency / FP32-issue bound like cfg2 (SURVEY H8); the mandated HBM fraction of the dominant "
                           "kernel, each kernel timed alone with CUDA events"),
    )
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_throughput(args.config)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
