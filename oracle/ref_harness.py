"""Run the UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py; or a checkout named by
DRB_REFERENCE_DIR / at /root/reference) on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `cpu_baseline` and `--impl reference` legs and tests/ use it; the
product never imports it.  What is timed is the reference's own public entry for the hot path,
`model_cl.RANSACLayer(opt).forward(points[N,4], weights[N], K1, K2, im1, im2)` (model_cl.py:160-256 ->
ransac.py:41-200: Gumbel top-5 sample, Nister five-point on every sample, MSAC score of every model against
every correspondence, arg-max, final refit), one pair after the other exactly as `DeepRansac_CLNet.forward`
does (model_cl.py:488-510), with `ransac_batch_size = max_iterations = K` so that one call scores K hypotheses
(SURVEY 8d "CPU reference timing").

Harness patches (none edits a reference file): an empty `h5py` module (feature_utils.py:7 imports it
for an unrelated loader) and `np.bool = bool` (loss.py:134 uses the alias numpy removed).

    python oracle/ref_harness.py --workload cfg2 --pairs 4 --hyps 1000 --corrs 2000 [--threads T] [--budget S]

prints one JSON line: hypotheses/s, threads, where the reference came from and whether its files hash to the
manifest.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_dir():
    """The directory the reference is imported from: $DRB_REFERENCE_DIR, else oracle/_ref, else /root/reference."""
    for cand in (os.environ.get("DRB_REFERENCE_DIR"), os.path.join(HERE, "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "ransac.py")):
            return cand
    return None


def setup_imports():
    """Make `import ransac, model_cl, loss, ...` resolve to the reference.  Returns its directory."""
    import numpy as np

    ref = reference_dir()
    if ref is None:
        raise ImportError("no reference: run `python oracle/make_ref.py` in the build container")
    if not hasattr(np, "bool"):
        np.bool = bool
    try:
        import h5py  # noqa: F401
    except ImportError:
        sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return ref


def provenance():
    ref = reference_dir()
    if ref is None:
        return dict(dir=None, verified=False)
    ok = None
    if os.path.exists(os.path.join(ref, "MANIFEST.json")):
        sys.path.insert(0, HERE)
        try:
            import make_ref

            ok = bool(make_ref.verify(ref))
        finally:
            sys.path.remove(HERE)
    return dict(dir=os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref,
                verified="sha256 of every file equals MANIFEST.json" if ok else ("checkout (no manifest)" if ok is None else False))


def make_opt(K, fmat=0, sampler=2, tr=0, threshold=0.75, device="cpu", precision=1):
    return types.SimpleNamespace(device=device, fmat=fmat, sampler=sampler, precision=precision, tr=tr,
                                 threshold=threshold, ransac_batch_size=int(K), weighted=0)


def make_layer(K, fmat=0, sampler=2, tr=0, max_iterations=None, precision=1):
    """The reference's RANSACLayer on the CPU with one chunk of K hypotheses per call (`precision` as `-pr`:
    1 = fp32, 2 = fp64, utils.py:42)."""
    setup_imports()
    import model_cl  # the reference's

    layer = model_cl.RANSACLayer(make_opt(K, fmat=fmat, sampler=sampler, tr=tr, precision=precision))
    layer.estimator.max_iterations = int(max_iterations or K)
    return layer


def intrinsics(focal=800.0):
    import torch

    K = torch.tensor([[focal, 0.0, 320.0], [0.0, focal, 240.0], [0.0, 0.0, 1.0]])
    return K, K.clone(), torch.tensor([480.0, 640.0]), torch.tensor([480.0, 640.0])


def time_layer(matches, weights, K, threads=None, budget_s=12.0, max_pairs=None, warmup_hyps=32):
    """hypotheses/s of the reference's RANSACLayer.forward over the pairs of `matches` [P,N,4] / `weights` [P,N]
    (stops after `budget_s` seconds; at least one pair)."""
    import torch

    if threads:
        torch.set_num_threads(int(threads))
    K1, K2, im1, im2 = intrinsics()
    with torch.no_grad():
        make_layer(warmup_hyps).forward(matches[0], weights[0], K1, K2, im1, im2)      # LAPACK / thread-pool warm-up
        layer = make_layer(K)
        done, total, per_pair = 0, 0.0, []
        for b in range(matches.shape[0] if max_pairs is None else min(max_pairs, matches.shape[0])):
            t0 = time.perf_counter()
            layer.forward(matches[b], weights[b], K1, K2, im1, im2)
            dt = time.perf_counter() - t0
            per_pair.append(dt)
            total += dt
            done += 1
            if total > budget_s:
                break
    return dict(value=done * K / total, pairs=done, seconds=total, per_pair_s=per_pair, threads=torch.get_num_threads())


def pick_threads(matches, weights, K=48, candidates=None):
    """The reference's best thread count on this host (its per-sample eigvals loop, nister.py:355-370, and the tiny
    batched LAPACK calls are slower with every core awake than with a few)."""
    import torch

    cores = os.cpu_count() or 1
    K1, K2, im1, im2 = intrinsics()
    best = (1, float("inf"))
    with torch.no_grad():
        for nt in sorted({1, 4, 8, 16, 32, cores} if candidates is None else set(candidates)):
            if nt > cores:
                continue
            torch.set_num_threads(nt)
            layer = make_layer(K)
            layer.forward(matches[0], weights[0], K1, K2, im1, im2)
            t0 = time.perf_counter()
            layer.forward(matches[0], weights[0], K1, K2, im1, im2)
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (nt, dt)
    torch.set_num_threads(best[0])
    return best[0], cores


def _gt_E_numpy(E):
    import numpy as np

    return np.asarray(E.detach().cpu().numpy(), dtype=np.float64)


def train_step(kind, K, points, weights, gt=None, focal=800.0):
    """One training step of the UNMODIFIED reference for ONE pair on the CPU (train.py:150-175 without the weight
    network): forward through RANSACLayer / RANSACLayer3D in train mode, the loss, and backward to `weights`.
        kind "e5":    5PC-E, MatchLoss(fmat=0) (loss.py:107-153; its GT-inlier mask via cv2.recoverPose on the host)
        kind "f8":    8PC-F (`-fmat 1 -sam 3`), MatchLoss(fmat=1); `points` are pixel coords / max(im) - centre / max(im)
        kind "rigid": RANSACLayer3D, its own mean squared residual (model_cl.py:516-595)
    Returns (loss value, d loss / d weights)."""
    import torch

    setup_imports()
    import loss as ref_loss
    import model_cl

    w = weights.clone().requires_grad_(True)
    K1, K2, im1, im2 = intrinsics(focal)
    if kind == "rigid":
        layer = model_cl.RANSACLayer3D(make_opt(K, tr=1, sampler=2))
        layer.estimator.max_iterations = int(K)
        _, l, _, _ = layer.forward(points, w, gt)
    else:
        fmat = 1 if kind == "f8" else 0
        layer = model_cl.RANSACLayer(make_opt(K, fmat=fmat, sampler=3 if fmat else 2, tr=1))
        layer.estimator.max_iterations = int(K)
        Es, _ = layer.forward(points, w, K1, K2, im1, im2, gt)
        gt_E = gt if not fmat else K2.T @ gt @ K1
        l = ref_loss.MatchLoss(fmat).forward([Es], _gt_E_numpy(gt_E)[None], [points[:, 0:2]], [points[:, 2:4]], [K1],
                                             [K2], [im1], [im2])
    l.backward()
    return float(l.detach()), w.grad


def stewenius_loop_body(matches, weights, K, threshold):
    """BASELINE cfg1: the reference's loop body (ransac.py:55-144) for one chunk of K hypotheses with its Stewenius
    estimator -- which cannot run as shipped (`self.device` is never set; D2: its estimate_model lacks the
    refit keywords, so only the loop body runs, not the whole __call__).  The harness sets `est.device`; nothing
    else is touched.  -> (best score, scores [K*10])."""
    import torch

    setup_imports()
    from estimators.essential_matrix_estimator_stewenius import EssentialMatrixEstimator
    from samplers.gumbel_sampler import GumbelSoftmaxSampler
    from scorings.msac_score import MSACScore

    est = EssentialMatrixEstimator("cpu")
    est.device = "cpu"
    smp = GumbelSoftmaxSampler(K, 5, device="cpu", data_type=torch.float32)
    samples, _ = smp.sample(weights)
    pts = matches.repeat([K, 1, 1]) * samples.unsqueeze(-1)
    minimal = pts[samples != 0].view(K, -1, 4)
    models = est.estimate_model(minimal)
    scores, masks = MSACScore("cpu").score(matches, models, threshold)
    return float(scores.max()), scores


def time_callable(fn, budget_s=12.0, max_reps=64, warmup=1):
    for _ in range(warmup):
        fn()
    times = []
    while len(times) < max_reps and (sum(times) < budget_s or not times):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--hyps", type=int, default=1000)
    ap.add_argument("--corrs", type=int, default=2000)
    ap.add_argument("--threads", type=int, default=0, help="0: fastest of a sweep")
    ap.add_argument("--budget", type=float, default=12.0)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    import torch

    import bench        # the workload generator only (bench.make_inputs: synthetic pairs, CPU tensors)

    matches, logits, _, _ = bench.make_inputs(args.pairs, args.corrs, seed=args.seed)
    threads, cores = (args.threads, os.cpu_count()) if args.threads else pick_threads(matches, logits)
    r = time_layer(matches, logits, args.hyps, threads=threads, budget_s=args.budget)
    r.update(unit="hypotheses/s", cores=threads, host_cores=cores, kind="reference", torch=torch.__version__,
             entry="model_cl.RANSACLayer.forward (unmodified reference)", **provenance())
    print(json.dumps(r))


if __name__ == "__main__":
    main()
