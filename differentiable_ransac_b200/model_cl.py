"""RANSACLayer / RANSACLayer3D / batch_episym -- host mirror of `model_cl.py:13-26, 160-256, 516-595`.

The nn.Module shells read the reference's `opt` namespace unchanged (`utils.py:7-83`); optional
engine knobs (`opt.seed`, `opt.adaptive`) default to the reference's behaviour.
`RANSACLayer.forward_batched` is the B-pairs-at-once entry that replaces the python loop of
`DeepRansac_CLNet.forward` (`model_cl.py:488-510`)."""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn

from . import engine
from .estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
from .estimators.fundamental_matrix_estimator import FundamentalMatrixEstimatorNew
from .estimators.rigid_transformation_SVD_based_solver import RigidTransformationSVDBasedSolver
from .loss import denormalize_pts
from .ransac import RANSAC, RANSAC3D, normalized_threshold
from .samplers.gumbel_sampler import GumbelSoftmaxSampler
from .scorings.msac_score import MSACScore


def batch_episym(x1, x2, F):
    """model_cl.py:13-26: x1, x2 [K,P,2] (the same P points for every model), F [K,3,3] -> [K,P].
    Unclamped distances; computed with torch ops (the fused, clamped, reduced version that the
    training loss uses is engine.EpisymLoss)."""
    one = x1.new_ones(*x1.shape[:2], 1)
    h1, h2 = torch.cat((x1, one), -1), torch.cat((x2, one), -1)
    Fx1 = torch.einsum("kij,kpj->kpi", F, h1)
    Ftx2 = torch.einsum("kji,kpj->kpi", F, h2)
    r = (h2 * Fx1).sum(-1)
    return r ** 2 * (1.0 / (Fx1[..., 0] ** 2 + Fx1[..., 1] ** 2 + 1e-15) + 1.0 / (Ftx2[..., 0] ** 2 + Ftx2[..., 1] ** 2 + 1e-15))


_PRECISION_NOTED = set()


def _dtype(opt):
    """`-pr` (utils.py:42): 0 = fp16, 1 = fp32, 2 = fp64 -- in the reference the dtype of the sampler's one-hot, which
    the rest of the chain follows by type promotion (test.py:13-18, nister.py:121-122).  `-pr 2` runs the test-mode
    five-point chain (solver, MSAC, arg-max, winner mask) in float64 on the device (csrc/fp64_path.cu,
    engine.ransac_e5_test_f64) when the adaptive exit and LO are off; every other chain computes in fp32 whatever `-pr`
    says (inputs converted on the way in, results returned in the requested dtype; refit, pose recovery and the
    backward adjoints are fp64 regardless).  Said once per process instead of silently."""
    p = getattr(opt, "precision", 1)
    if p != 1 and p not in _PRECISION_NOTED:
        import warnings

        _PRECISION_NOTED.add(p)
        warnings.warn(f"-pr {p}: only the test-mode five-point chain without adaptive exit / LO has a float64 device path; "
                      f"the other CUDA chains compute in fp32 (refit, pose recovery and backward adjoints in fp64) and "
                      f"return {'float64' if p == 2 else 'float16'} tensors")
    return {2: torch.float64, 0: torch.float16}.get(p, torch.float32)


class RANSACLayer(nn.Module):
    def __init__(self, opt, **kwargs):
        super().__init__(**kwargs)
        self.opt = opt
        solver = FundamentalMatrixEstimatorNew(opt.device, opt.weighted) if opt.fmat else EssentialMatrixEstimatorNister(opt.device)
        if opt.sampler in (0, 1):
            raise NotImplementedError("sampler ids 0/1 cannot run through the reference driver either (SURVEY 2.1 #5)")
        s = solver.sample_size if opt.sampler == 2 else 8
        sampler = GumbelSoftmaxSampler(opt.ransac_batch_size, s, device=opt.device, data_type=_dtype(opt),
                                       seed=getattr(opt, "seed", 0))
        if opt.fmat:
            max_iters = 1000 if opt.tr else 5000            # model_cl.py:213-219
        else:
            max_iters = 100 if opt.tr else 5000
        self.estimator = RANSAC(solver, sampler, MSACScore(opt.device), max_iterations=max_iters, fmat=opt.fmat,
                                train=opt.tr, ransac_batch_size=opt.ransac_batch_size, sampler_id=opt.sampler,
                                weighted=opt.weighted, threshold=opt.threshold,
                                adaptive=getattr(opt, "adaptive", True), final_refit=getattr(opt, "final_refit", True))

    def _points(self, points, im_size1, im_size2):
        pts = points.clone()
        if self.opt.fmat:
            pts[..., 0:2] = denormalize_pts(points[..., 0:2].clone(), im_size1)
            pts[..., 2:4] = denormalize_pts(points[..., 2:4].clone(), im_size2)
        return pts

    def forward(self, points, weights, K1, K2, im_size1, im_size2, ground_truth=None):
        pts = self._points(points, im_size1, im_size2)
        torch.cuda.synchronize()
        t0 = time.time()
        models, _, _, _ = self.estimator(pts, weights.reshape(-1), K1, K2, ground_truth)
        torch.cuda.synchronize()
        dt = time.time() - t0
        Es = torch.cat(list(models.values())) if self.opt.tr else models
        if Es.dim() == 3:
            Es = Es[~torch.isnan(Es).flatten(1).any(1)]
        return Es, dt

    def thresholds(self, K1, K2, B, device):
        """ransac.py:49-53 for B pairs at once: threshold / ((K1[0,0] + K1[1,1] + K1[0,0] + K2[1,1]) / 4), computed
        where K1 / K2 live (no device->host sync when they are on the GPU, as test.py:31 puts them)."""
        drv = self.estimator
        if drv.fmat:
            return torch.full((B,), float(drv.threshold), dtype=torch.float32, device=device)
        K1, K2 = torch.as_tensor(K1), torch.as_tensor(K2)
        mult = (K1[..., 0, 0] + K1[..., 1, 1] + K1[..., 0, 0] + K2[..., 1, 1]) / 4
        return (float(drv.threshold) / mult).to(dtype=torch.float32).expand(B).to(device)

    def forward_batched(self, points, weights, K1, K2, im_size1=None, im_size2=None, ground_truth=None, K=None):
        """points [B,N,4], weights [B,N], K1/K2 [B,3,3] -> list_B[Es_b] (train: [K_b,3,3]; test: [3,3]),
        with ONE launch per stage for the whole batch.  HOST tensors in test mode go through the pipelined
        service (`submit` / `collect`); the winners' inlier masks and scores of the last call are kept in
        `self.last_masks` / `self.last_scores`."""
        B = points.shape[0]
        if not self.opt.tr and points.device.type == "cpu" and self._service_ok():
            Es, self.last_masks, self.last_scores = self.collect(self.submit(points, weights, K1, K2, K=K))
            return Es
        pts = points
        if self.opt.fmat:
            pts = torch.stack([self._points(points[b], im_size1[b], im_size2[b]) for b in range(B)])
        drv = self.estimator
        smp = drv.sampler
        if self.opt.tr:
            Kt = K or drv._chunks() * drv.ransac_batch_size
            if smp.num_samples == 8:
                models, valid = engine.HypothesizeF8.apply(pts, weights, Kt, smp.tau, None, smp.seed, smp._next_offset())
            else:
                models, valid = engine.HypothesizeE5.apply(pts, weights, ground_truth.float(), Kt, smp.tau, None,
                                                           smp.seed, smp._next_offset(), True)
            return [models[b][valid[b]] for b in range(B)]
        out = drv.batched_test(pts, weights, self.thresholds(K1, K2, B, points.device), K)
        self.last_masks, self.last_scores = out["mask"], out["best_score"]
        return list(out["best_model"].unbind(0))

    # -- pipelined entry for batches that live in HOST memory (what a DataLoader hands to model_cl.py:488-510) ----
    def _service_ok(self):
        """The pipelined service runs the hot path proper -- sample -> five-point -> MSAC -> arg-max + winner mask,
        every hypothesis scored -- i.e. test mode, essential matrix, no adaptive exit / LO / final refit
        (`opt.adaptive = False`, `opt.final_refit = False`: SURVEY 8d's cfg2 semantics)."""
        drv = self.estimator
        return (not self.opt.tr and not drv.fmat and drv.sampler.num_samples == 5 and not drv.adaptive
                and not drv.final_refit and not drv.lo and torch.cuda.is_available())

    def submit(self, points, weights, K1, K2, K=None, slots=3, graph=True):
        """Enqueue one batch of HOST tensors (points [B,N,4], weights [B,N], K1/K2 [B,3,3] or [3,3]) and return a
        ticket; nothing blocks unless `slots` batches are already in flight.  Pinned tensors are copied
        asynchronously straight from where they are.  `collect(ticket)` returns the batch's results."""
        if not self._service_ok():
            raise NotImplementedError("submit/collect serve test-mode essential-matrix layers with opt.adaptive = "
                                      "False and opt.final_refit = False; use forward_batched otherwise")
        drv = self.estimator
        B, N, _ = points.shape
        Kh = int(K or drv._chunks() * drv.ransac_batch_size)
        key = (B, N, Kh, int(slots), bool(graph))
        if getattr(self, "_svc_key", None) != key:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._svc = engine.E5TestService(B, N, Kh, dev, slots=slots, seed=drv.sampler.seed, graph=graph,
                                             want_mask=True)
            self._svc_key = key
            self._thr_pinned = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(int(slots))]
            self._thr_np = [t.numpy() for t in self._thr_pinned]
        svc = self._svc
        slot = svc.step % svc.slots
        if svc.busy[slot]:
            svc.done[slot].synchronize()           # its results were not collected: they are overwritten
        thr = self._thr_pinned[slot]
        if (torch.is_tensor(K1) and torch.is_tensor(K2) and K1.device.type == "cpu" and K2.device.type == "cpu"
                and K1.dtype == torch.float32 and K2.dtype == torch.float32):
            # ransac.py:49-53 on the host without a dozen framework dispatches (the step has ~100 us of host time)
            k1, k2 = K1.numpy(), K2.numpy()
            np.divide(4.0 * float(drv.threshold), k1[..., 0, 0] + k1[..., 1, 1] + k1[..., 0, 0] + k2[..., 1, 1],
                      out=self._thr_np[slot], casting="unsafe")
        else:
            thr.copy_(self.thresholds(K1, K2, B, "cpu"))
        return svc.submit(slot, host=(points, weights, thr))

    def collect(self, ticket):
        """-> (list_B[E_b [3,3]], masks [B,N] bool, scores [B]) of the batch `submit` returned `ticket` for: views
        of the slot's pinned host buffers, valid until the slot is submitted again (`slots` submits later)."""
        return self._svc.result_views(ticket)


class RANSACLayer3D(nn.Module):
    def __init__(self, opt, **kwargs):
        super().__init__(**kwargs)
        self.opt = opt
        solver = RigidTransformationSVDBasedSolver()
        sampler = GumbelSoftmaxSampler(opt.ransac_batch_size, 3 if opt.sampler != 3 else 8, device=opt.device,
                                       data_type=_dtype(opt), seed=getattr(opt, "seed", 0))
        self.estimator = RANSAC3D(solver, sampler, MSACScore(opt.device), max_iterations=1000, fmat=opt.fmat,
                                  train=opt.tr, ransac_batch_size=opt.ransac_batch_size, sampler_id=opt.sampler,
                                  weighted=opt.weighted, threshold=opt.threshold)

    def forward(self, points, weights, ground_truth=None):
        t0 = time.time()
        models, residuals, avg_residuals, _, _ = self.estimator(points, weights.reshape(-1), ground_truth)
        dt = time.time() - t0
        Es = torch.cat(list(models.values()))
        loss = torch.cat(list(residuals.values()))
        avg = sum(avg_residuals.values()) / len(avg_residuals)
        return Es, loss.mean(), avg, dt
