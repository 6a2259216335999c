"""Oracle: the RANSAC driver loop body (test and train branches) and RANSAC3D's
train branch.

Restates `ransac.py:49-144` (threshold normalisation with the K1[0,0]-twice
quirk of :52, chunked sample -> solve -> score -> argmax, train-mode
closest-to-GT selection :87-96) and `ransac.py:352-382` (3-D train branch),
with the Gumbel noise injected per chunk.  `full_test_driver` adds what follows the
loop body in test mode: adaptive exit `:134-142, :202-215`,
local optimisation `:217-257` (lo = 1, 2) and the final refit `:148-195`.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import torch

from . import fundamental, nister, rigid, sampler, scoring


def normalized_threshold(threshold: float, K1: torch.Tensor, K2: torch.Tensor, fmat: bool) -> float:
    """ransac.py:49-53 (note K1[0,0] appears twice)."""
    if fmat:
        return threshold
    return float(threshold / ((K1[0, 0] + K1[1, 1] + K1[0, 0] + K2[1, 1]) / 4))


def solve(minimal: torch.Tensor, solver: str):
    if solver == "nister":
        return nister.five_point(minimal), 10
    if solver == "f8":
        return fundamental.eight_point(minimal), 1
    raise ValueError(solver)


def test_loop(matches, logits, noises, threshold, solver="nister", sample_size=5, tau=1.0):
    """ransac.py:55-144, test branch, no adaptive exit: one chunk per entry of
    `noises` (each [K_c, N]).  Returns dict(best_model, best_mask, best_score,
    best_chunk, best_idx (index into the chunk's compacted model list),
    scores (list per chunk), models (list per chunk))."""
    best = dict(best_score=0, best_model=None, best_mask=None, best_chunk=-1, best_idx=-1,
                scores=[], models=[], samples=[])
    for ci, G in enumerate(noises):
        ret, _, idx = sampler.sample(logits, G, sample_size, tau)
        minimal = sampler.gather_minimal(matches, ret)
        models, _ = solve(minimal, solver)
        scores, masks = scoring.msac_score(matches, models, threshold)
        bi = int(torch.argmax(scores))
        best["scores"].append(scores)
        best["models"].append(models)
        best["samples"].append(idx)
        if scores[bi] > best["best_score"] or ci == 0:
            best.update(best_score=scores[bi], best_mask=masks[bi], best_model=models[bi], best_chunk=ci,
                        best_idx=bi)
    return best


def train_select(models: torch.Tensor, gt_model: torch.Tensor, n_slots: int) -> torch.Tensor:
    """ransac.py:87-96: per sample keep the slot closest to GT in raw Frobenius
    norm (sign-sensitive, as the reference)."""
    d = torch.norm(models - gt_model, dim=(1, 2)).view(-1, n_slots)
    pick = torch.argmin(d, dim=-1)
    return models.view(-1, n_slots, 3, 3)[torch.arange(pick.shape[0]), pick]


def train_loop(matches, logits, noises, gt_model, solver="nister", sample_size=5, tau=1.0):
    """ransac.py:78-108: collect one model per sample and chunk, NaN-filtered."""
    out = []
    for G in noises:
        ret, _, _ = sampler.sample(logits, G, sample_size, tau)
        minimal = sampler.gather_minimal(matches, ret)
        models, n_slots = solve(minimal, solver)
        chosen = models if n_slots == 1 else train_select(models, gt_model, n_slots)
        ok = ~torch.isnan(chosen).flatten(1).any(1)
        out.append(chosen[ok])
    return torch.cat(out)


def rigid_train_loop(points, logits, noises, flag=True, tau=1.0):
    """ransac.py:352-382: per chunk (models [K_c,4,4], residual sums [K_c],
    mean residual scalar)."""
    models, res, mean_res = [], [], []
    for G in noises:
        ret, _, _ = sampler.sample(logits, G, 3, tau)
        minimal = sampler.gather_minimal(points, ret)
        m, _, _, _ = rigid.estimate(minimal, flag=flag)
        r, mr, _ = scoring.rigid_squared_residual(points[:, :3], points[:, 3:], m[:, :3, :].transpose(-1, -2))
        models.append(m)
        res.append(r)
        mean_res.append(mr)
    return models, res, mean_res


def _nonminimal(points, fmat, weights=None):
    """What `estimator.estimate_model` does for n > sample_size rows: fundamental_matrix_estimator.py:169-175
    (normalise + eight-point on all rows) or, without pymagsac, nister.py:51-65 (five-point system on the
    four smallest right singular vectors of A^T A)."""
    if fmat:
        return fundamental.eight_point(points, weights)
    return nister.five_point(points)


def adaptive_iteration_number(inlier_number, point_number, confidence, sample_size, max_iterations, eps=1e-5):
    """ransac.py:202-215."""
    import math
    ratio = float(inlier_number) / point_number
    if 1.0 - ratio ** sample_size >= 1.0 - eps:
        return max_iterations
    return max(0.0, math.log10(1.0 - confidence) / math.log10(1 - ratio ** sample_size + eps))


def local_optimization(matches, best_score, best_mask, best_model, threshold, fmat, lo, lo_iters):
    """ransac.py:217-257 for lo in {1, 2}: refit on the current inliers, keep while the score does not drop."""
    iters = lo_iters if lo == 2 else 1
    for _ in range(iters):
        points = matches[best_mask].unsqueeze(0)
        models = _nonminimal(points, fmat)
        scores, masks = scoring.msac_score(matches, models, threshold)
        bi = int(torch.argmax(scores))
        if scores[bi] >= best_score:
            best_score, best_mask, best_model = scores[bi], masks[bi], models[bi]
        else:
            break
    return best_score, best_mask, best_model


def full_test_driver(matches, logits, noises, K1, K2, threshold, fmat=False, sample_size=5, tau=1.0,
                     confidence=0.999, lo=0, lo_iters=64, weighted=False, estimator_sample_size=None):
    """`RANSAC.__call__` in test mode, ransac.py:41-200: chunked loop with adaptive exit, optional LO on every
    improvement, final refit (8-point on the inliers / 5-point on all points in fp64), MSAC re-score, keep
    the refit only if it scores higher.  -> (best_model, best_mask, best_score, iterations).
    The exponent of the adaptive budget is the ESTIMATOR's sample size (ransac.py:207,214 read
    `self.estimator.sample_size`): 5 for the five-point, 7 for FundamentalMatrixEstimatorNew
    (fundamental_matrix_estimator.py:163) even when the sampler draws 8 points (`-fmat 1 -sam 3`)."""
    if estimator_sample_size is None:
        estimator_sample_size = 7 if fmat else 5
    rbs = noises[0].shape[0]
    max_iterations = rbs * len(noises)
    thr = normalized_threshold(threshold, K1, K2, fmat)
    N = matches.shape[0]
    solver = "f8" if fmat else "nister"
    best_score, best_mask, best_model = 0, None, None
    iterations, max_iters, ci = 0, max_iterations, 0
    while iterations < max_iters:
        ret, y_soft, _ = sampler.sample(logits, noises[ci], sample_size, tau)
        minimal = sampler.gather_minimal(matches, ret)
        models, _ = solve(minimal, solver)
        scores, masks = scoring.msac_score(matches, models, thr)
        bi = int(torch.argmax(scores))
        if scores[bi] > best_score or iterations == 0:
            best_score, best_mask, best_model = scores[bi], masks[bi], models[bi]
            if lo:
                best_score, best_mask, best_model = local_optimization(matches, best_score, best_mask, best_model,
                                                                       thr, fmat, lo, lo_iters)
            max_iters = min(max_iterations, adaptive_iteration_number(int(best_mask.sum()), N, confidence,
                                                                      estimator_sample_size, max_iterations))
        iterations += rbs
        ci += 1
    if fmat:
        w = y_soft[0][best_mask][None] if weighted else None
        cand = fundamental.eight_point(matches[best_mask].unsqueeze(0), w)
    else:
        cand = nister.five_point(matches.unsqueeze(0).double()).to(matches.dtype)
    scores, _ = scoring.msac_score(matches, cand, thr)
    if scores.max() > best_score:
        bi = int(torch.argmax(scores))
        best_model, best_score = cand[bi], scores[bi]
    return best_model, best_mask, best_score, iterations
