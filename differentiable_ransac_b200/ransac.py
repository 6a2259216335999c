"""RANSAC / RANSAC3D -- host mirror of the reference drivers (`ransac.py:6-200`, `:303-450`).

Same constructor, same `__call__` signature and return tuple.  Inside, the reference's
`while iterations < max_iters` loop of torch ops is replaced by launches of the CUDA path:
every chunk is one `sample -> solve -> score -> arg-max` pass for `ransac_batch_size`
hypotheses (test mode), or ONE pass for all chunks at once (train mode, where nothing
depends on the previous chunk).  `batched_call` does the same for B pairs in one go.
"""
from __future__ import annotations

import math

import torch

from . import engine, ops


def normalized_threshold(threshold, K1, K2, fmat):
    """ransac.py:49-53, including the K1[0,0]-twice quirk."""
    if fmat:
        return float(threshold)
    return float(threshold) / float((K1[0, 0] + K1[1, 1] + K1[0, 0] + K2[1, 1]) / 4)


def threshold_tensor(threshold, K1, K2, fmat, device):
    """The same value as a [1] fp32 tensor on `device`, computed where K1 / K2 live: no device->host sync when
    they are CUDA tensors (test.py:31 moves them to the GPU)."""
    if fmat or not torch.is_tensor(K1):
        return torch.tensor([normalized_threshold(threshold, K1, K2, fmat)], dtype=torch.float32, device=device)
    mult = (K1[0, 0] + K1[1, 1] + K1[0, 0] + K2[1, 1]) / 4
    return (float(threshold) / mult).reshape(1).to(device=device, dtype=torch.float32)


class RANSAC(object):
    def __init__(self, estimator, sampler, scoring, fmat=False, train=False, ransac_batch_size=64, sampler_id=0,
                 weighted=0, threshold=1e-3, confidence=0.999, max_iterations=5000, lo=0, lo_iters=64, eps=1e-5,
                 adaptive=True, final_refit=True):
        self.estimator, self.sampler, self.scoring = estimator, sampler, scoring
        self.lo, self.lo_iters = lo, lo_iters
        self.fmat, self.train = fmat, train
        self.ransac_batch_size = ransac_batch_size
        self.sampler_id = sampler_id
        self.weighted = weighted
        self.threshold = threshold
        self.confidence = confidence
        self.max_iterations = max_iterations
        self.eps = eps
        self.adaptive = adaptive          # reference behaviour: early exit by inlier count (one host sync per chunk)
        self.final_refit = final_refit
        if sampler_id not in (2, 3):
            raise NotImplementedError("only the Gumbel samplers (ids 2, 3) run in the reference (SURVEY 2.1 #5)")

    # -- helpers -------------------------------------------------------------------------------
    @property
    def sample_size(self):
        return self.sampler.num_samples

    def _chunks(self):
        return int(math.ceil(self.max_iterations / self.ransac_batch_size))

    def adaptive_iteration_number(self, inlier_number, point_number, confidence):
        """ransac.py:202-215."""
        inlier_ratio = float(inlier_number) / point_number
        probability = 1.0 - inlier_ratio ** self.estimator.sample_size
        if probability >= 1.0 - self.eps:
            return self.max_iterations
        return max(0.0, math.log10(1.0 - confidence) / math.log10(1 - inlier_ratio ** self.estimator.sample_size + self.eps))

    # -- the driver ------------------------------------------------------------------------------
    def plugins_are_native(self):
        """True when sampler, estimator and scoring are exactly this package's classes: then `__call__` runs the
        fused CUDA path for them.  Anything else -- a user's estimator or scoring, or a subclass that overrides a
        method -- is HONOURED: the chunked loop is run through the objects' own `sample` / `estimate_model` /
        `score` calls, as the reference does (ransac.py:63-76, 111)."""
        from .estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
        from .estimators.essential_matrix_estimator_stewenius import EssentialMatrixEstimator
        from .estimators.fundamental_matrix_estimator import FundamentalMatrixEstimatorNew
        from .samplers.gumbel_sampler import GumbelSoftmaxSampler
        from .scorings.msac_score import MSACScore

        return (type(self.estimator) in (EssentialMatrixEstimatorNister, EssentialMatrixEstimator,
                                         FundamentalMatrixEstimatorNew)
                and type(self.sampler) is GumbelSoftmaxSampler and type(self.scoring) is MSACScore)

    def _plugin_call(self, matches, logits, K1, K2, gt_model):
        """The reference's loop (ransac.py:55-200) driven through the plugin objects' own methods, for plugins that
        are not this package's classes.  One chunk per trip: `sampler.sample(logits)` -> dense straight-through
        one-hot [K,N]; gather; `estimator.estimate_model(minimal[K,s,D] (, weights))`; then either the train-mode
        selection (closest slot to `gt_model`, NaN rows dropped) or `scoring.score(matches, models, thr)`, arg-max,
        strict improvement and the adaptive iteration budget.  No local optimisation here (it refits through this
        package's kernels, which a foreign estimator does not use): lo != 0 raises."""
        if self.lo:
            raise NotImplementedError("local optimisation runs on this package's solvers; with a plugin estimator / "
                                      "scoring use lo=0")
        thr = normalized_threshold(self.threshold, K1, K2, self.fmat)
        rbs, N, D = self.ransac_batch_size, matches.shape[0], matches.shape[-1]
        iterations, budget = 0, self.max_iterations
        best_score, best_mask, best_model, soft = 0, [], [], None
        collected = {}
        while iterations < budget:
            ret, soft = self.sampler.sample(logits)
            picked = ret != 0
            minimal = (matches.unsqueeze(0) * ret.unsqueeze(-1))[picked].view(rbs, -1, D)
            if minimal.shape[1] == 0:
                continue
            if self.weighted:
                est = self.estimator.estimate_model(minimal, soft[picked].view(rbs, -1))
            else:
                est = self.estimator.estimate_model(minimal)
            if self.train:
                if est is None or est.shape[0] == 0:
                    continue
                if self.sampler.num_samples == 8:
                    chosen = est
                else:
                    slots = 4 if self.fmat else 10
                    grouped = est.view(-1, slots, 3, 3)
                    pick = (grouped - gt_model).flatten(2).norm(dim=-1).argmin(dim=-1)
                    chosen = grouped[torch.arange(grouped.shape[0], device=est.device), pick]
                collected[iterations] = chosen[~torch.isnan(chosen).flatten(1).any(1)]
            else:
                scores, masks = self.scoring.score(matches, est, thr)
                i = torch.argmax(scores)
                if iterations == 0 or scores[i] > best_score:
                    best_score, best_mask, best_model = scores[i], masks[i], est[i]
                    if self.adaptive:
                        budget = min(self.max_iterations,
                                     self.adaptive_iteration_number(int(best_mask.sum()), N, self.confidence))
            iterations += rbs
        if self.train:
            return collected, best_mask, best_score, iterations
        if self.final_refit:                      # ransac.py:148-185, the same calls on the plugin estimator
            inl = best_mask.nonzero(as_tuple=True)
            if self.fmat:
                args = (soft[0, inl[0]],) if self.weighted else ()
                est = self.estimator.estimate_model(matches[inl].unsqueeze(0), *args)
            else:
                est = self.estimator.estimate_model(matches.unsqueeze(0).double(), K1=K1.detach().cpu().numpy(),
                                                    K2=K2.detach().cpu().numpy(),
                                                    inlier_indices=inl[0].cpu().numpy().astype("uint64"),
                                                    best_model=best_model.detach().cpu().numpy().T,
                                                    unnormalzied_threshold=0.75, best_score=best_score)
            if est is None or est.shape[0] == 0:
                best_model = torch.eye(3, device=matches.device, dtype=matches.dtype)
            else:
                est = est.to(matches.dtype)
                scores, _ = self.scoring.score(matches, est, thr)
                if scores.max() > best_score:
                    best_model, best_score = est[torch.argmax(scores)], scores.max()
        if not getattr(self.scoring, "provides_inliers", True):
            best_model, best_mask = self.scoring.get_inliers(matches, best_model.unsqueeze(0), self.estimator,
                                                             threshold=thr)
        return best_model, best_mask, best_score, iterations

    def __call__(self, matches, logits, K1, K2, gt_model):
        if not self.plugins_are_native():
            return self._plugin_call(matches, logits, K1, K2, gt_model)
        if self.train:
            return self._train(matches, logits, gt_model)
        return self._test(matches, logits, threshold_tensor(self.threshold, K1, K2, self.fmat, matches.device))

    def _train(self, matches, logits, gt_model):
        """ransac.py:78-108: every chunk's chosen models, NaN-free, keyed by the iteration count."""
        rbs = self.ransac_batch_size
        K = self._chunks() * rbs
        noise = self.sampler.injected_noise
        if noise is not None:
            noise = noise.reshape(1, -1, noise.shape[-1])
        seed, off = self.sampler.seed, self.sampler._next_offset()
        m, lg = matches[None], logits[None]
        if self.sample_size == 8:
            models, valid = engine.HypothesizeF8.apply(m, lg, K, self.sampler.tau, noise, seed, off)
        elif self.sample_size == 5:
            models, valid = engine.HypothesizeE5.apply(m, lg, gt_model[None].float(), K, self.sampler.tau, noise, seed,
                                                       off, True)
        else:
            raise NotImplementedError("train mode supports the 5-point and the 8-point samplers")
        out = {}
        for c in range(self._chunks()):
            sl = slice(c * rbs, (c + 1) * rbs)
            out[c * rbs] = models[0, sl][valid[0, sl]]
        return out, [], 0, K

    def _run(self):
        run = {5: engine.ransac_e5_test, 7: engine.ransac_f7_test, 8: engine.ransac_f8_test}.get(self.sample_size)
        if run is None:
            raise NotImplementedError("test mode supports the 5-, 7- and 8-point samplers")
        return run

    def _loop(self, m, lg, thr, noise):
        """ransac.py:55-144 for B pairs without LO, no host sync: every chunk in one pass, the loop's bookkeeping
        (and its adaptive exit) replayed on the device.  noise: [B, chunks*rbs, N] or None."""
        rbs, K = self.ransac_batch_size, self._chunks() * self.ransac_batch_size
        off = self.sampler._next_offset()
        if self.adaptive:
            return engine.ransac_test_adaptive(m, lg, rbs, self.max_iterations, thr, self.sample_size,
                                               self.confidence, self.eps, self.sampler.tau, noise, self.sampler.seed,
                                               off, adaptive_exponent=self.estimator.sample_size)
        run = self._run()
        if self.sample_size == 5 and getattr(self.sampler, "dtype", torch.float32) == torch.float64:
            run = engine.ransac_e5_test_f64        # `-pr 2`: solver, MSAC, arg-max and mask in float64 (fp64_path.cu)
        out = run(m, lg, K, thr, self.sampler.tau, noise, self.sampler.seed, off)
        out["iterations"] = torch.full((m.shape[0],), K, dtype=torch.int32, device=m.device)
        return out

    def _loop_with_lo(self, m, lg, thr, noise):
        """The same loop when LO runs after every improvement (ransac.py:122-132): the next comparison depends on
        the optimised score, so the chunks are sequential -- one pair, one host sync per chunk, like the
        reference."""
        if self.lo not in (1, 2):
            raise NotImplementedError("lo=3 needs the reference's UniformSampler, which cannot run (SURVEY 2.1 #5)")
        rbs, N = self.ransac_batch_size, m.shape[1]
        best, iterations, max_iters, chunk = None, 0, self.max_iterations, 0
        while iterations < max_iters:
            nz = None if noise is None else noise[:, chunk * rbs:(chunk + 1) * rbs]
            out = self._run()(m, lg, rbs, thr, self.sampler.tau, nz, self.sampler.seed, self.sampler._next_offset())
            if best is None or bool(out["best_score"][0] > best["best_score"][0]):     # ransac.py:116
                best = engine.local_optimization(m, out, thr, self.fmat, self.lo_iters if self.lo == 2 else 1)
                if self.adaptive:
                    max_iters = min(self.max_iterations,
                                    self.adaptive_iteration_number(int(best["ninl"][0]), N, self.confidence))
            iterations += rbs
            chunk += 1
        best["iterations"] = torch.full((1,), iterations, dtype=torch.int32, device=m.device)
        return best

    def _test(self, matches, logits, thr):
        f64 = getattr(self.sampler, "dtype", torch.float32) == torch.float64      # `-pr 2`
        m, lg = (matches[None].double() if f64 else matches[None].float()), logits[None].float()
        noise = self.sampler.injected_noise
        if noise is not None:
            noise = noise.reshape(1, -1, noise.shape[-1])
        best = self._loop_with_lo(m, lg, thr, noise) if self.lo else self._loop(m, lg, thr, noise)
        iterations = int(best["iterations"][0])
        if self.final_refit:
            # ransac.py:148-185.  `weighted` hands the eight-point the soft one-hot of the last chunk's first
            # sample (:153); that row exists only when the noise is injected, otherwise softmax(logits / tau).
            w = None
            if self.fmat and self.weighted:
                key = lg if noise is None else lg + noise[:, iterations - self.ransac_batch_size].to(lg.dtype)
                w = torch.softmax(key / self.sampler.tau, dim=-1)
            best = engine.final_refit(m, best, thr, self.fmat, w)
        best_model, best_mask, best_score = best["best_model"][0], best["mask"][0], best["best_score"][0]
        return best_model.to(matches.dtype), best_mask, best_score, iterations

    # -- B pairs at once (replaces the python loop of model_cl.py:488-510) ---------------------------
    def batched_test(self, matches, logits, thresholds, K=None):
        """matches [B,N,4], logits [B,N], thresholds [B] -> engine result dict (winner, inlier mask, score,
        `iterations` [B]) for all pairs at once: the chunked loop with its adaptive exit replayed on the device,
        then the final refit.  With lo = 0 this is, per pair, exactly what `__call__` returns.  With lo in (1, 2)
        it is NOT: `__call__` optimises after every improvement inside the loop (ransac.py:122-132, sequential per
        pair), here local optimisation runs ONCE on the loop's winner.  `K` overrides max_iterations."""
        keep = self.max_iterations
        if K:
            self.max_iterations = int(K)
        try:
            out = self._loop(matches, logits, thresholds, None)
        finally:
            self.max_iterations = keep
        if self.lo in (1, 2):
            # batched LO runs once on the loop's winner (inside the loop it is sequential per pair: `__call__`)
            out = engine.local_optimization(matches, out, thresholds, self.fmat, self.lo_iters if self.lo == 2 else 1)
        if self.final_refit:
            out = engine.final_refit(matches, out, thresholds, self.fmat)
        return out


class RANSAC3D(RANSAC):
    """`ransac.py:303-450`.  Only the train branch of the reference runs (its test branch reads
    undefined variables), so that is the branch provided."""

    def __call__(self, matches, logits, gt_model, valid=False):
        if valid or not self.train:
            raise NotImplementedError("RANSAC3D test mode is broken in the reference (ransac.py:384-396)")
        rbs = self.ransac_batch_size
        chunks = self._chunks()
        K = chunks * rbs
        noise = self.sampler.injected_noise
        if noise is not None:
            noise = noise.reshape(1, -1, noise.shape[-1])
        models, ok = engine.HypothesizeRigid.apply(matches[None], logits[None], K, True, self.sampler.tau, noise,
                                                   self.sampler.seed, self.sampler._next_offset())
        res = engine.RigidResidual.apply(matches[None], models)[0]
        N = matches.shape[0]
        out_m, out_r, out_mean = {}, {}, {}
        for c in range(chunks):
            sl = slice(c * rbs, (c + 1) * rbs)
            keep = ok[0, sl]
            # the reference drops NaN models inside estimate_model, before squared_residual
            # (rigid...solver.py:60-74): residuals and their mean cover the valid models only
            out_m[c * rbs] = models[0, sl][keep]
            out_r[c * rbs] = res[sl][keep]
            out_mean[c * rbs] = res[sl][keep].sum() / (keep.sum().clamp_min(1) * N)
        return out_m, out_r, out_mean, 0, K
