"""Batched hypothesize-and-score pipelines and their autograd wrappers.

This is the B200-side replacement of the reference's per-pair Python loop
(`model_cl.py:488-510`) around `RANSAC.__call__` (`ransac.py:41-200`): all B pairs and
all K hypotheses go through one launch per stage -- sample -> minimal solve ->
residual/score -> arg-max -- with nothing of size K x N ever materialised.

Test mode  : `ransac_e5_test`, `ransac_f8_test`          (ransac.py:109-142 without early exit)
Train mode : `HypothesizeE5`, `HypothesizeF8`, `HypothesizeRigid` (ransac.py:78-108 / :352-382)
Loss       : `EpisymLoss` (loss.py:138-151), `RigidResidual` (rigid...solver.py:76-89)
"""
from __future__ import annotations

import torch

import os

from . import ops

# Scorer of the pipelined service (E5TestService with more than one slot):
#   "auto"     (default) the fastest exact-operand tensor-core variant that verifies itself on this device --
#              "tc_bf16p" (0.105 ms at cfg2), then "tc_bf16" (0.122), then "tc_tf32" -- else "block"; see
#              service_scorer().  Every variant is pinned to the fp64 oracle at 1e-4 (BF16) / 5e-4 (TF32) by
#              tests/test_gpu_score_tc.py (profiles/r2_tc_oracle_parity.log)
#   "tc_tf32"  tensor cores, two TF32 words per operand (csrc/score_tc.cu): measured on B200 at cfg2 0.175 ms per
#              pipelined batch against 0.254 ms with "block" (profiles/r1_notes.md); per-model scores within 1.4e-4
#              relative of the FP32 kernels (the winner's score: ~1e-6), so the winner can differ from the FP32
#              kernels' only between near-ties
#   "tc_bf16"  three BF16 words per operand: exact operands, fp32-level scores (5e-6 against fp64 on the host model),
#              same time
#   "tc_bf16p" the same with ONE reciprocal per pair of neighbouring models (halves the SFU work)
#   "tc_bf16p_s" tc_bf16p built slim (128 registers, three stages): 3 % slower alone, but it leaves room on every SM for a
#              five-point CTA of the next batch -- the pipelined step is 2.6 % faster (0.1395 vs 0.1432 ms at cfg2)
#   "block"    FP32, one CTA per 32 models; bit-identical to what `ransac_e5_test(scorer="block")` returns
#   "stream"   FP32 work queue
SERVICE_SCORER = os.environ.get("DRB_SERVICE_SCORER", "auto")


def _noise_args(noise, seed, offset):
    return dict(noise=noise, seed=int(seed), offset=int(offset))


_TC_CHECKED = {}


def tc_scorer_agrees(device, scorer="tc_tf32", tol=5e-4):
    """One-off check (cached per device and scorer) that a tensor-core scorer reproduces the FP32 kernel on a
    problem built to expose a broken operand split: 256 random models x 4000 random correspondences with a tight
    threshold (1e-3), where losing the second-order partial products of the BF16 split moves scores by 5e-3 and a
    correct split by < 1e-4 (profiles/r1_notes.md).  Scores within `tol` relative (floor 1.0) of the FP32 block
    kernel's."""
    key = (str(device), scorer)
    if key in _TC_CHECKED:
        return _TC_CHECKED[key]
    ok = False
    try:
        gen = torch.Generator().manual_seed(1234)
        B, M, N = 2, 256, 4000
        matches = (torch.rand(B, N, 4, generator=gen) - 0.5).to(device)
        models = torch.randn(B, M, 3, 3, generator=gen)
        models = (models / models.flatten(-2).norm(dim=-1)[..., None, None]).to(device)
        thr = torch.full((B,), 1e-3).to(device)
        s_ref, _ = ops.score_msac(matches, models, thr, kernel="block")
        s_tc, _ = ops.score_msac(matches, models, thr, kernel=scorer)
        rel = ((s_tc - s_ref).abs() / s_ref.clamp_min(1.0)).max()
        ok = bool(torch.isfinite(s_tc).all()) and float(rel) < tol and float(s_ref.max()) > 1.0
    except Exception as e:          # a scorer that cannot even run is a failed check
        import warnings

        warnings.warn(f"tensor-core scorer {scorer!r} could not be checked: {e}")
        ok = False
    _TC_CHECKED[key] = ok
    return ok


def service_scorer(device, B, requested=None):
    """The scorer a pipelined service uses: `requested` (or SERVICE_SCORER) if it names one; "auto" -> the first of
    "tc_bf16p_s", "tc_bf16p", "tc_bf16", "tc_tf32" that agrees with the FP32 kernel on this device (tc_scorer_agrees), else "block".  A
    tensor-core scorer named by SERVICE_SCORER is checked the same way and replaced by "block" (with a warning) if
    it fails; one passed explicitly by the caller is taken as is.  drb_score_msac_tc takes at most 1024 pairs."""
    if requested is not None:
        return requested
    name = SERVICE_SCORER
    if name != "auto" and not name.startswith("tc"):
        return name
    if int(B) > 1024:
        return "block"
    for cand in (("tc_bf16p_s", "tc_bf16p", "tc_bf16", "tc_tf32") if name == "auto" else (name,)):
        if tc_scorer_agrees(device, cand):
            return cand
    import warnings

    warnings.warn("no tensor-core scorer agrees with the FP32 kernel on this device: the service scores with the "
                  "FP32 block kernel")
    return "block"


# ---- test mode ---------------------------------------------------------------------------------
def _draw(logits, K, s, tau, noise, seed, offset, sampler, offset_dev=None):
    """Test-mode sampling: injected noise -> exact Gumbel keys (bit-exact with the reference);
    otherwise the set sampler (same distribution, no K x N noise) unless `sampler="gumbel"`."""
    if noise is None and sampler == "sets":
        return ops.sample_sets(logits, K, s, int(seed), int(offset), offset_dev)
    if offset_dev is not None:
        raise ops._lib.DrbError("a device-side stream offset needs the set sampler")
    return ops.sample(logits, K, s, tau, **_noise_args(noise, seed, offset))[0]


def ransac_e5_test(matches, logits, K, thr, tau=1.0, noise=None, seed=0, offset=0, want_scores=False,
                   sampler="sets", offset_dev=None, scorer=None):
    """matches [B,N,4], logits [B,N], thr [B] (normalised threshold, ransac.py:49-53).
    Returns dict(best_model [B,3,3], best_hyp [B], best_slot [B], best_score [B], mask [B,N] bool,
    ninl [B], idx [B,K,5], models [B,K,10,3,3], nsol [B,K] (, scores [B,K*10] in compact order,
    cids)).

    scorer: "auto" (the tensor-core scorer the pipelined service would pick on this device),
    "stream" (default; persistent work-queue kernel: lowest latency for one call, and it keeps every SM
    busy when there are few models) or "block" (one CTA per 32 models: its CTAs retire one by one, which lets
    the kernels of an independent call on another stream move in -- what pipelined callers want, see
    E5TestService)."""
    B = matches.shape[0]
    if scorer == "auto":        # what the pipelined service would pick on this device (tensor cores when they verify)
        scorer = service_scorer(matches.device, B, None if SERVICE_SCORER == "auto" else SERVICE_SCORER)
    idx = _draw(logits, K, 5, tau, noise, seed, offset, sampler, offset_dev)
    best0, cc0 = ops.zeroed_counters(B, matches.device)
    models, nsol, cm, cid, cc = ops.solve_e5(matches, idx, compact=True, ccount=cc0)
    scores, best = ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=want_scores, best=best0,
                                  kernel=scorer)
    best_id, best_score, best_model, mask, ninl = ops.best_finalize(matches, models.reshape(B, -1, 9), best, thr)
    # best_id = hypothesis * 10 + slot; mask is 0/1 bytes, reinterpreted (not copied) as bool
    out = dict(best_model=best_model, best_id=best_id, best_score=best_score, mask=mask.view(torch.bool), ninl=ninl,
               idx=idx, models=models, nsol=nsol)
    if want_scores:
        out.update(scores=scores, cids=cid, ccount=cc,
                   best_hyp=torch.div(best_id, ops.E5_SLOTS, rounding_mode="floor"), best_slot=best_id % ops.E5_SLOTS)
    return out


def ransac_e5_test_f64(matches, logits, K, thr, tau=1.0, noise=None, seed=0, offset=0, want_scores=False,
                       sampler="sets"):
    """`ransac_e5_test` with the solver, MSAC, arg-max and winner mask in float64 -- what the reference computes under
    `-pr 2` (utils.py:42, model_cl.py:164-170: the sampler's one-hot carries the dtype and the chain follows by type
    promotion).  The samples are drawn exactly as in the fp32 path (same sampler, same noise).  Returns the same
    dict with float64 models and scores (csrc/fp64_path.cu; built for results, not for the roofline)."""
    idx = _draw(ops._f32(logits), K, 5, tau, noise, seed, offset, sampler)
    models, nsol = ops.solve_e5_f64(matches, idx)
    scores = ops.score_msac_f64(matches, models, thr, nsol)
    best_id, best_score, best_model, mask, ninl = ops.best_finalize_f64(matches, models, scores, thr)
    out = dict(best_model=best_model, best_id=best_id, best_score=best_score, mask=mask.view(torch.bool), ninl=ninl,
               idx=idx, models=models, nsol=nsol,
               best_hyp=torch.div(best_id, ops.E5_SLOTS, rounding_mode="floor"), best_slot=best_id % ops.E5_SLOTS)
    if want_scores:
        out["scores"] = scores
    return out


def ransac_f8_test(matches, logits, K, thr, tau=1.0, noise=None, seed=0, offset=0, want_scores=False,
                   sampler="sets"):
    idx = _draw(logits, K, 8, tau, noise, seed, offset, sampler)
    models, valid = ops.solve_f8(matches, idx)
    scores, best = ops.score_msac(matches, models, thr, want_scores=want_scores)
    best_id, best_score, best_model, mask, ninl = ops.best_finalize(matches, models, best, thr)
    out = dict(best_model=best_model, best_id=best_id, best_hyp=best_id, best_score=best_score, mask=mask.bool(),
               ninl=ninl, idx=idx, models=models, valid=valid.bool())
    if want_scores:
        out["scores"] = scores
    return out


def ransac_f7_test(matches, logits, K, thr, tau=1.0, noise=None, seed=0, offset=0, want_scores=False,
                   sampler="sets"):
    """Fundamental matrix from 7-point samples (`-fmat 1 -sam 2` in the reference, whose own 7-point is broken:
    SURVEY D4).  Up to three models per sample; slots without a real root are NaN for the scorer, which never
    lets a NaN score win."""
    idx = _draw(logits, K, 7, tau, noise, seed, offset, sampler)
    models, nsol = ops.solve_f7(matches, idx)                                   # [B,K,3,3,3], [B,K]
    B = matches.shape[0]
    live = torch.arange(ops.F7_SLOTS, device=models.device)[None, None, :] < nsol[..., None]
    scored = torch.where(live[..., None, None], models, torch.full_like(models, float("nan"))).reshape(B, -1, 9)
    scores, best = ops.score_msac(matches, scored, thr, want_scores=want_scores)
    best_id, best_score, best_model, mask, ninl = ops.best_finalize(matches, models.reshape(B, -1, 9), best, thr)
    out = dict(best_model=best_model, best_id=best_id, best_score=best_score, mask=mask.view(torch.bool), ninl=ninl,
               idx=idx, models=models, nsol=nsol)
    if want_scores:
        out.update(scores=scores, best_hyp=torch.div(best_id, ops.F7_SLOTS, rounding_mode="floor"))
    return out


# ---- the chunked loop with adaptive exit, no host sync -----------------------
def ransac_test_adaptive(matches, logits, rbs, max_iterations, thr, sample_size=5, confidence=0.999, eps=1e-5,
                         tau=1.0, noise=None, seed=0, offset=0, sampler="sets", adaptive_exponent=None):
    """What `RANSAC.__call__` returns in test mode with lo = 0 (ransac.py:55-144) for B pairs at once: all
    C = ceil(max_iterations / rbs) chunks go through sample -> solve -> score in one pass, then
    `drb_adaptive_select` replays the loop's bookkeeping (per-chunk arg-max, strict improvement, adaptive
    iteration budget from the winner's inlier count) on the device.  `noise` (optional) is [B, C*rbs, N], chunk c
    = rows [c*rbs, (c+1)*rbs).  -> dict(best_model, best_id, best_score, mask, ninl, iterations [B], ...).
    `sample_size` picks the solver (the SAMPLER's draw size); `adaptive_exponent` is the exponent of the iteration
    budget, the ESTIMATOR's `sample_size` in the reference (ransac.py:207,214) -- 7 for the fundamental-matrix
    estimator even with 8-point samples; defaults to `sample_size`."""
    B = matches.shape[0]
    C = -(-int(max_iterations) // int(rbs))
    K = C * int(rbs)
    idx = _draw(logits, K, sample_size, tau, noise, seed, offset, sampler)
    if sample_size == 5:
        slots = ops.E5_SLOTS
        best0, cc0 = ops.zeroed_counters(B, matches.device)
        models, nsol, cm, cid, cc = ops.solve_e5(matches, idx, compact=True, ccount=cc0)
        scores, _ = ops.score_msac(matches, cm, thr, count=cc, ids=cid, want_scores=True, best=best0)
        extra = dict(nsol=nsol)
    elif sample_size == 7:
        slots = ops.F7_SLOTS
        models, nsol = ops.solve_f7(matches, idx)
        live = torch.arange(slots, device=models.device)[None, None, :] < nsol[..., None]
        scored = torch.where(live[..., None, None], models, torch.full_like(models, float("nan"))).reshape(B, -1, 9)
        scores, _ = ops.score_msac(matches, scored, thr, want_scores=True)
        cc = cid = None
        extra = dict(nsol=nsol)
    elif sample_size == 8:
        slots = 1
        models, valid = ops.solve_f8(matches, idx)
        scores, _ = ops.score_msac(matches, models, thr, want_scores=True)
        cc = cid = None
        extra = dict(valid=valid.bool())
    else:
        raise NotImplementedError("test mode supports the 5-, 7- and 8-point samplers")
    dense = models.reshape(B, -1, 9)
    best, its, chunk_best, chunk_ninl = ops.adaptive_select(matches, dense, scores, thr, int(rbs) * slots, rbs,
                                                            max_iterations, int(adaptive_exponent or sample_size),
                                                            confidence, eps, cc, cid)
    best_id, best_score, best_model, mask, ninl = ops.best_finalize(matches, dense, best, thr)
    out = dict(best_model=best_model, best_id=best_id, best_score=best_score, mask=mask.view(torch.bool), ninl=ninl,
               iterations=its, idx=idx, models=models, chunk_ninl=chunk_ninl)
    out.update(extra)
    return out


# ---- after the loop: local optimisation and the final refit -----------------
def _fit_and_score(matches, mask, thr, fmat, weights=None):
    """Non-minimal fit on the selected correspondences of every pair, scored on all correspondences:
    -> (score [B], model [B,3,3], inlier mask [B,N] bool, ninl [B]) of the best candidate per pair
    (score 0 / identity when the fit produced nothing)."""
    fit = ops.refit_f8 if fmat else ops.refit_e5
    models, nsol = fit(matches, mask, weights)
    _, packed = ops.score_msac(matches, models, thr, count=nsol, want_scores=False)
    _, score, model, inl, ninl = ops.best_finalize(matches, models, packed, thr)
    return score, model, inl.view(torch.bool), ninl


def local_optimization(matches, state, thr, fmat, iters):
    """`RANSAC.localOptimization` for lo = 1 (iters = 1) and lo = 2 (iters = lo_iters), ransac.py:217-257, for
    B pairs at once: refit on the current inliers, accept while the score does not drop (`>=`, :250), stop a
    pair at its first rejection.  `state` = dict(best_score [B], best_model [B,3,3], mask [B,N], ninl [B]);
    returns the updated dict.  One host sync per iteration (the reference syncs on the same comparison)."""
    score, model, mask, ninl = state["best_score"], state["best_model"], state["mask"], state["ninl"]
    active = torch.ones_like(score, dtype=torch.bool)
    for _ in range(int(iters)):
        s2, m2, k2, n2 = _fit_and_score(matches, mask, thr, fmat)
        take = active & (s2 >= score)
        score = torch.where(take, s2, score)
        model = torch.where(take[:, None, None], m2, model)
        mask = torch.where(take[:, None], k2, mask)
        ninl = torch.where(take, n2, ninl)
        active = take
        if not bool(active.any()):
            break
    out = dict(state)
    out.update(best_score=score, best_model=model, mask=mask, ninl=ninl)
    return out


def final_refit(matches, state, thr, fmat, weights=None):
    """ransac.py:148-185: the eight-point on the winner's inliers (`fmat`), or the five-point system on ALL
    correspondences (what `estimate_model` does without pymagsac, nister.py:51-65; the reference runs it in
    fp64, so does the kernel).  The candidate replaces the winner only if it scores strictly higher (:181);
    the inlier mask is NOT recomputed (the reference keeps the loop's mask)."""
    fit = ops.refit_f8 if fmat else ops.refit_e5
    models, nsol = fit(matches, state["mask"] if fmat else None, weights)
    _, packed = ops.score_msac(matches, models, thr, count=nsol, want_scores=False)
    _, s2, m2, _, _ = ops.best_finalize(matches, models, packed, thr, want_mask=False)
    take = s2 > state["best_score"]
    out = dict(state)
    out.update(best_score=torch.where(take, s2, state["best_score"]),
               best_model=torch.where(take[:, None, None], m2, state["best_model"]), refit_taken=take)
    return out


# ---- train mode --------------------------------------------------------------------------------
class _HypothesizeBase(torch.autograd.Function):
    """sample -> minimal solve (-> slot selection); backward: solver IFT adjoint ->
    gather backward -> straight-through sampler backward.  Gradients reach `matches`
    and `logits` (what train.py / train_ransac_loftr.py need)."""

    @staticmethod
    def _backward_common(ctx, g_pts):
        matches, logits, idx, lse, sel_key = ctx.saved_tensors[:5]
        g_sel, gm = ops.gather_backward(matches, idx, g_pts, want_grad_matches=ctx.needs_input_grad[0])
        gl = None
        if ctx.needs_input_grad[1]:
            gl = ops.sample_backward(logits, idx, lse, sel_key, g_sel, ctx.tau, ctx.noise, ctx.seed, ctx.offset)
        return gm, gl


class HypothesizeE5(_HypothesizeBase):
    @staticmethod
    def forward(ctx, matches, logits, gt, K, tau=1.0, noise=None, seed=0, offset=0, sign_invariant=True):
        idx, lse, sel_key, _ = ops.sample(logits, K, 5, tau, noise, seed, offset, want_lse=True)
        sel, chosen, _, _ = ops.solve_e5_select(matches, idx, gt, sign_invariant)      # solve + slot choice, one launch
        ctx.save_for_backward(ops._f32(matches), ops._f32(logits), idx, lse, sel_key, chosen, sel)
        ctx.tau, ctx.noise, ctx.seed, ctx.offset = float(tau), noise, int(seed), int(offset)
        valid = sel >= 0
        ctx.mark_non_differentiable(valid)
        return chosen, valid

    @staticmethod
    def backward(ctx, g_chosen, _g_valid):
        matches, logits, idx, lse, sel_key, chosen, sel = ctx.saved_tensors
        g_pts = ops.solve_e5_backward_chosen(matches, idx, chosen, sel, g_chosen.reshape(*sel.shape, 9))
        gm, gl = _HypothesizeBase._backward_common(ctx, g_pts)
        return gm, gl, None, None, None, None, None, None, None


class HypothesizeF8(_HypothesizeBase):
    @staticmethod
    def forward(ctx, matches, logits, K, tau=1.0, noise=None, seed=0, offset=0):
        idx, lse, sel_key, _ = ops.sample(logits, K, 8, tau, noise, seed, offset, want_lse=True)
        models, valid = ops.solve_f8(matches, idx)
        ctx.save_for_backward(ops._f32(matches), ops._f32(logits), idx, lse, sel_key, models)
        ctx.tau, ctx.noise, ctx.seed, ctx.offset = float(tau), noise, int(seed), int(offset)
        valid = valid.bool()
        ctx.mark_non_differentiable(valid)
        return models, valid

    @staticmethod
    def backward(ctx, g_models, _g_valid):
        matches, logits, idx, lse, sel_key, models = ctx.saved_tensors
        g_pts = ops.solve_f8_backward(matches, idx, g_models.reshape(*idx.shape[:2], 9), models)
        gm, gl = _HypothesizeBase._backward_common(ctx, g_pts)
        return gm, gl, None, None, None, None, None


class HypothesizeRigid(_HypothesizeBase):
    @staticmethod
    def forward(ctx, points, logits, K, flag=True, tau=1.0, noise=None, seed=0, offset=0):
        idx, lse, sel_key, _ = ops.sample(logits, K, 3, tau, noise, seed, offset, want_lse=True)
        models, valid = ops.solve_rigid3(points, idx, flag)
        ctx.save_for_backward(ops._f32(points), ops._f32(logits), idx, lse, sel_key)
        ctx.tau, ctx.noise, ctx.seed, ctx.offset, ctx.flag = float(tau), noise, int(seed), int(offset), bool(flag)
        valid = valid.bool()
        ctx.mark_non_differentiable(valid)
        return models, valid

    @staticmethod
    def backward(ctx, g_models, _g_valid):
        points, logits, idx, lse, sel_key = ctx.saved_tensors
        g_pts = ops.solve_rigid3_backward(points, idx, g_models.reshape(*idx.shape[:2], 16), ctx.flag)
        gm, gl = _HypothesizeBase._backward_common(ctx, g_pts)
        return gm, gl, None, None, None, None, None, None


# ---- losses ------------------------------------------------------------------------------------
class EpisymLoss(torch.autograd.Function):
    """row_sum[b,k] = sum_p min(episym(pts[b,p], models[b,k]), 1)   (loss.py:138-144).
    Differentiable w.r.t. models only (the correspondences are data in train.py)."""

    @staticmethod
    def forward(ctx, pts, models, npts=None, mvalid=None):
        out = ops.episym_forward(pts, models, npts, mvalid)
        ctx.save_for_backward(ops._f32(pts), ops._f32(models))
        ctx.npts, ctx.mvalid = npts, mvalid
        return out

    @staticmethod
    def backward(ctx, g_row):
        pts, models = ctx.saved_tensors
        g = ops.episym_backward(pts, models, g_row, ctx.npts, ctx.mvalid)
        return None, g.reshape(models.shape), None, None


class RigidResidual(torch.autograd.Function):
    """res_sum[b,k] = sum_n ||q_n - (R p_n + t)||^2   (rigid...solver.py:76-89)."""

    @staticmethod
    def forward(ctx, points, models):
        res, _ = ops.rigid_residual_forward(points, models, want_ninl=False)
        ctx.save_for_backward(ops._f32(points), ops._f32(models))
        return res

    @staticmethod
    def backward(ctx, g_res):
        points, models = ctx.saved_tensors
        return None, ops.rigid_residual_backward(points, models, g_res).reshape(models.shape)


def match_loss(models, valid, pts, npts=None):
    """Mean over valid models and points of min(episym, 1): the MatchLoss value of
    loss.py:138-151 for every pair.  models [B,K,3,3], valid [B,K], pts [B,P,4] (the GT-inlier
    correspondences, zero-padded to P with npts[b] real ones).  Returns [B]."""
    B, P, _ = pts.shape
    row = EpisymLoss.apply(pts, models, npts, valid)
    n = (npts if npts is not None else torch.full((B,), P, device=pts.device)).to(row.dtype)
    v = valid.to(row.dtype)
    return (row * v).sum(1) / (v.sum(1).clamp_min(1.0) * n.clamp_min(1.0))


# ---- one training step without autograd bookkeeping, optionally ONE CUDA graph ------------------------
class TrainStep:
    """Forward + loss + backward of the hot path for one batch as a fixed sequence of C-ABI launches on static
    buffers -- what `train.py:150-175` does through autograd (`HypothesizeE5/F8/Rigid` -> `match_loss` /
    `RigidResidual` -> `.backward()`), minus the ~25 small torch ops, the graph building and the per-step
    allocations that made the cfg5 step host-bound (0.6-0.7 ms in round 1 for ~0.3 ms of kernels).

        kind "e5":    sample 5 -> Nister five-point -> slot closest to the GT model -> min(episym, 1) over the
                      GT-inlier correspondences (loss.py:138-151), mean over valid hypotheses and pairs
        kind "f8":    sample 8 -> normalised eight-point -> the same loss
        kind "rigid": sample 3 -> rigid 3-point -> sum of squared residuals over all points
                      (rigid...solver.py:76-89), mean over valid hypotheses, points and pairs

    `run(...)` copies the batch into the static buffers and replays; afterwards `loss` [1], `loss_pairs` [B],
    `grad_logits` [B,N] (= d loss / d logits, what CLNet's backward consumes, train.py:171) and, when asked,
    `grad_matches` hold the step's results.  The Philox stream position lives on the device and advances inside
    the graph: every replay draws fresh Gumbel noise, and the backward regenerates exactly the forward's.
    Equal to the autograd path launch for launch (tests/test_gpu_train_step.py)."""

    def __init__(self, kind, B, N, K, device, P=None, seed=0, tau=1.0, graph=True, want_grad_matches=False,
                 sign_invariant=True, flag=True):
        if kind not in ("e5", "f8", "rigid"):
            raise ValueError(kind)
        self.kind, self.B, self.N, self.K, self.tau = kind, int(B), int(N), int(K), float(tau)
        self.seed, self.graph_mode = int(seed), bool(graph)
        self.want_gm, self.sign_invariant, self.flag = bool(want_grad_matches), bool(sign_invariant), bool(flag)
        self.device = torch.device(device)
        D = 6 if kind == "rigid" else 4
        z = dict(dtype=torch.float32, device=self.device)
        self.matches = torch.zeros(B, N, D, **z)
        self.logits = torch.zeros(B, N, **z)
        if kind != "rigid":
            self.P = int(P if P is not None else N)
            self.pts = torch.zeros(B, self.P, 4, **z)          # GT-inlier correspondences, zero-padded
            self.npts = torch.zeros(B, dtype=torch.int32, device=self.device)
        if kind == "e5":
            self.gt = torch.zeros(B, 3, 3, **z)
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.g = None
        self.loss = self.loss_pairs = self.grad_logits = self.grad_matches = None

    def _body(self):
        B, K, N = self.B, self.K, self.N
        s = {"e5": 5, "f8": 8, "rigid": 3}[self.kind]
        m, lg, ctr = self.matches, self.logits, self.counter
        idx, lse, sel_key, _ = ops.sample(lg, K, s, self.tau, None, self.seed, 0, want_lse=True, offset_dev=ctr)
        if self.kind == "e5":
            sel, used, _, _ = ops.solve_e5_select(m, idx, self.gt, self.sign_invariant)
            valid = sel >= 0
        elif self.kind == "f8":
            used, valid = ops.solve_f8(m, idx)
            valid = valid.bool()
        else:
            used, valid = ops.solve_rigid3(m, idx, self.flag)
            valid = valid.bool()
        # d mean_b(loss_b) / d row depends only on which models are valid: known BEFORE the loss pass, so the loss and
        # its gradient with respect to the models come out of ONE pass over the points
        v = valid.to(torch.float32)
        if self.kind == "rigid":
            denom = v.sum(1).clamp_min(1.0) * float(N)
            g_row = v / (denom[:, None] * float(B))
            row, g_used = ops.rigid_residual_forward_backward(
                m, torch.where(valid[..., None, None], used, torch.zeros_like(used)), g_row)   # invalid models carry NaN
            g_pts = ops.solve_rigid3_backward(m, idx, g_used.reshape(B, K, 16), self.flag)
        else:
            denom = v.sum(1).clamp_min(1.0) * self.npts.to(torch.float32).clamp_min(1.0)
            g_row = v / (denom[:, None] * float(B))
            row, g_used = ops.episym_forward_backward(self.pts, used, g_row, self.npts, valid)
            if self.kind == "e5":
                g_pts = ops.solve_e5_backward_chosen(m, idx, used, sel, g_used.reshape(B, K, 9))
            else:
                g_pts = ops.solve_f8_backward(m, idx, g_used.reshape(B, K, 9), used)
        loss_pairs = (row * v).sum(1) / denom
        g_sel, gm = ops.gather_backward(m, idx, g_pts, want_grad_matches=self.want_gm)
        gl = ops.sample_backward(lg, idx, lse, sel_key, g_sel, self.tau, None, self.seed, 0, offset_dev=ctr)
        ctr.add_(1)
        return loss_pairs.mean().reshape(1), loss_pairs, gl, gm

    def _capture(self):
        st = self.stream
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            self._body()                                   # eager warm-up (lazy attributes, allocator pools)
        st.synchronize()
        self.counter.zero_()
        self.g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g, stream=st):
            self.loss, self.loss_pairs, self.grad_logits, self.grad_matches = self._body()

    def load(self, matches, logits, gt=None, pts=None, npts=None):
        """Copy a batch (device or pinned-host tensors) into the static buffers, on the current stream."""
        self.matches.copy_(matches, non_blocking=True)
        self.logits.copy_(logits, non_blocking=True)
        if gt is not None:
            self.gt.copy_(gt, non_blocking=True)
        if pts is not None:
            self.pts.copy_(pts, non_blocking=True)
            self.npts.copy_(npts, non_blocking=True)

    def run(self, matches=None, logits=None, gt=None, pts=None, npts=None):
        """One step on the current stream; returns (loss [1], grad_logits [B,N]) -- static tensors, overwritten by
        the next `run`."""
        if matches is not None:
            self.load(matches, logits, gt, pts, npts)
        if self.graph_mode:
            if self.g is None:
                self._capture()
            self.g.replay()
        else:
            self.loss, self.loss_pairs, self.grad_logits, self.grad_matches = self._body()
        return self.loss, self.grad_logits


# ---- host-buffer service: test-mode batches from pinned host memory, pipelined over streams -----------
class E5TestService:
    """Steady-state test-mode service for batches that live in HOST memory (the reference's
    `model_cl.py:488-510` loop receives its pairs from a DataLoader).  `submit()` enqueues, for one batch of
    B pairs: one packed host->device copy (matches | logits | thr), `ransac_e5_test`, and one packed
    device->host copy of (best model, best id, best score, #inliers); `result()` waits for that batch and
    returns host tensors.  Nothing blocks in `submit()` unless every slot is in flight.

    `slots` batches are in flight at once, each on its own compute stream, with the copies on a separate copy
    stream.  Batches are independent, so the latency-bound 5-point kernel of batch i+1 fills the issue slots
    the FMA-bound scoring kernel of batch i leaves idle (measured on B200, cfg2: 0.309 -> 0.256 ms per batch
    with two slots, profiles/r1_notes.md)."""

    def __init__(self, B, N, K, device, slots=2, seed=0, graph=False, host_io=True, scorer=None, want_mask=False):
        self.B, self.N, self.K, self.seed = int(B), int(N), int(K), int(seed)
        # one slot: whatever a single call uses (ops default); several: service_scorer() unless the caller says
        self.scorer = scorer if (scorer is not None or int(slots) <= 1) else service_scorer(device, B)
        self.graph = bool(graph)
        # host_io=False: batches are already on the device -- submit(slot, packed=<device tensor>) copies the
        # packed batch into the slot (device to device) and results stay on the device (`dev_out[slot]`)
        self.host_io = bool(host_io)
        # want_mask: the winner's inlier mask [B,N] (what `RANSAC.__call__` returns as best_mask, ransac.py:200)
        # is part of every batch's results: a second, B*N-byte copy beside the packed floats
        self.want_mask = bool(want_mask)
        self.device = torch.device(device)
        self.slots = int(slots)
        self.n_in = B * N * 4 + B * N + B
        self.n_out = B * 9 + 3 * B
        self.host_in = [torch.empty(self.n_in, dtype=torch.float32).pin_memory() for _ in range(self.slots)]
        self.host_out = [torch.empty(self.n_out, dtype=torch.float32).pin_memory() for _ in range(self.slots)]
        self.dev_in = [torch.empty(self.n_in, dtype=torch.float32, device=self.device) for _ in range(self.slots)]
        self.dev_out = [torch.empty(self.n_out, dtype=torch.float32, device=self.device) for _ in range(self.slots)]
        if self.want_mask:
            self.host_mask = [torch.empty(B, N, dtype=torch.uint8).pin_memory() for _ in range(self.slots)]
            self.dev_mask = [torch.empty(B, N, dtype=torch.uint8, device=self.device) for _ in range(self.slots)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.compute = [torch.cuda.Stream(device=self.device) for _ in range(self.slots)]
        self.copied = [torch.cuda.Event() for _ in range(self.slots)]
        self.consumed = [torch.cuda.Event() for _ in range(self.slots)]
        self.done = [torch.cuda.Event() for _ in range(self.slots)]
        self.busy = [False] * self.slots
        self.step = 0
        self.h2d_bytes = self.n_in * 4
        self.d2h_bytes = self.n_out * 4 + (B * N if self.want_mask else 0)
        # graph mode: one CUDA graph per slot (copy-in, four kernels, copy-out), replayed by submit(); the Philox
        # stream position of the slot's next batch lives on the device and is advanced inside the graph, so
        # batch i draws with offset i exactly as in eager mode when the slots are used round-robin
        self.counters = [torch.full((1,), s, dtype=torch.int64, device=self.device) for s in range(self.slots)]
        self.graphs = [None] * self.slots
        # Host-side cost of a step (the device needs 0.14 ms at cfg2; Python has to stay well below that): the views
        # of the slot buffers are made once, the input copies go through one foreign call each (drb_copy_h2d_async)
        # and `result_views` hands out the same objects every time -- they alias the slot's pinned buffers.
        self._in_views = [(b_[: B * N * 4].view(B, N, 4), b_[B * N * 4: B * N * 5].view(B, N), b_[B * N * 5:])
                          for b_ in self.dev_in]
        self._in_ptrs = [tuple(v.data_ptr() for v in views) for views in self._in_views]
        self._in_bytes = (B * N * 16, B * N * 4, B * 4)
        self._copy = ops._lib.load().drb_copy_h2d_async
        self._views = [None] * self.slots

    def _body(self, slot, offset, offset_dev):
        B, N = self.B, self.N
        buf = self.dev_in[slot]
        o = ransac_e5_test(buf[: B * N * 4].view(B, N, 4), buf[B * N * 4: B * N * 5].view(B, N), self.K,
                           buf[B * N * 5:], seed=self.seed, offset=offset, offset_dev=offset_dev,
                           scorer=self.scorer)
        return o, torch.cat((o["best_model"].flatten(), o["best_id"].float(), o["best_score"], o["ninl"].float()))

    def _copy_out(self, slot, o, packed):
        """Results of the slot's batch -> its output buffers (pinned host memory when host_io)."""
        (self.host_out if self.host_io else self.dev_out)[slot].copy_(packed, non_blocking=True)
        if self.want_mask:
            (self.host_mask if self.host_io else self.dev_mask)[slot].copy_(o["mask"].view(torch.uint8), non_blocking=True)

    def _capture(self, slot):
        ks = self.compute[slot]
        ks.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ks):
            self._body(slot, 0, self.counters[slot])     # eager warm-up: lazy kernel attributes, allocator pools
        ks.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ks):
            # the host->device copy stays OUTSIDE the graph (its source may be the caller's own pinned tensors)
            o, packed = self._body(slot, 0, self.counters[slot])
            self.counters[slot].add_(self.slots)
            self._copy_out(slot, o, packed)
        self.graphs[slot] = g

    def stage(self, slot, matches, logits, thr):
        """Pack one batch of host tensors into the slot's pinned staging buffer (a host-side memcpy; callers
        whose tensors are pinned already hand them to `submit(host=...)` instead and skip this)."""
        B, N = self.B, self.N
        buf = self.host_in[slot]
        buf[: B * N * 4].copy_(matches.reshape(-1))
        buf[B * N * 4: B * N * 5].copy_(logits.reshape(-1))
        buf[B * N * 5:].copy_(thr.reshape(-1))

    def _h2d(self, slot, host):
        """Enqueue the slot's host->device copy on the current stream: from the staging buffer, or straight from the
        caller's (matches [B,N,4], logits [B,N], thr [B]) host tensors."""
        B, N = self.B, self.N
        buf = self.dev_in[slot]
        if host is None:
            buf.copy_(self.host_in[slot], non_blocking=True)
            return
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for src, dst, ptr, nbytes in zip(host, self._in_views[slot], self._in_ptrs[slot], self._in_bytes):
            if (src.dtype == torch.float32 and src.device.type == "cpu" and src.is_contiguous()
                    and src.numel() * 4 == nbytes):
                ops.check(self._copy(ptr, src.data_ptr(), nbytes, stream), "drb_copy_h2d_async")
            else:                                   # another dtype / layout / a device tensor: the framework's copy
                dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def submit(self, slot=None, packed=None, host=None):
        """Enqueue one batch; returns the slot.  host_io: the batch staged in `host_in[slot]`, or `host` =
        (matches, logits, thr) host tensors copied directly (pinned: asynchronously).  host_io=False: `packed`, one
        device tensor (matches | logits | thr, `n_in` floats)."""
        if slot is None:
            slot = self.step % self.slots
        if self.busy[slot] and self.host_io:
            self.done[slot].synchronize()          # the slot's previous results must have been collected
        cs, ks = self.copy_stream, self.compute[slot]
        if not self.host_io:
            if self.graph and self.graphs[slot] is None:
                self._capture(slot)
            # `packed` was produced on the caller's current stream: the slot's stream waits for it, and the
            # allocator must not hand its block to anyone while the copy is pending.  The same wait orders this
            # batch behind whatever the caller enqueued on its stream to read the slot's previous results
            # (`dev_out[slot]`); a consumer on any OTHER stream must `join()` before the slot comes round again.
            ks.wait_stream(torch.cuda.current_stream(self.device))
            packed.record_stream(ks)
            with torch.cuda.stream(ks):
                self.dev_in[slot].copy_(packed, non_blocking=True)
                if self.graph:
                    self.graphs[slot].replay()
                else:
                    o, out = self._body(slot, self.step, None)
                    self._copy_out(slot, o, out)
                self.done[slot].record(ks)
            self.busy[slot] = True
            self.step += 1
            return slot
        if self.graph:
            if self.graphs[slot] is None:
                self._capture(slot)
            with torch.cuda.stream(ks):            # in order on the slot's stream: copy-in, graph (kernels + copy-out)
                self._h2d(slot, host)
                self.graphs[slot].replay()
                self.done[slot].record(ks)
            self.busy[slot] = True
            self.step += 1
            return slot
        cs.wait_event(self.consumed[slot])         # the previous batch of this slot has read its inputs
        with torch.cuda.stream(cs):
            self._h2d(slot, host)
            self.copied[slot].record(cs)
        ks.wait_event(self.copied[slot])
        with torch.cuda.stream(ks):
            o, packed = self._body(slot, self.step, None)
            self.consumed[slot].record(ks)
            self._copy_out(slot, o, packed)
            self.done[slot].record(ks)
        self.busy[slot] = True
        self.step += 1
        return slot

    def result(self, slot):
        """Wait for the slot's batch; -> dict of tensors: views of the slot's pinned output buffer (host_io) or
        of its device output buffer."""
        self.done[slot].synchronize()
        self.busy[slot] = False
        B = self.B
        out = (self.host_out if self.host_io else self.dev_out)[slot]
        res = dict(best_model=out[: 9 * B].view(B, 3, 3), best_id=out[9 * B: 10 * B].to(torch.int32),
                   best_score=out[10 * B: 11 * B], ninl=out[11 * B: 12 * B].to(torch.int32))
        if self.want_mask:
            res["mask"] = (self.host_mask if self.host_io else self.dev_mask)[slot].view(torch.bool)
        return res

    def result_views(self, slot):
        """`result(slot)` for callers in a hurry (RANSACLayer.collect): waits for the slot's batch and returns
        (list_B[model [3,3]], mask [B,N] bool | None, score [B]) -- the SAME view objects on every call, aliasing the
        slot's pinned output buffers (valid until the slot is submitted again)."""
        self.done[slot].synchronize()
        self.busy[slot] = False
        v = self._views[slot]
        if v is None:
            B = self.B
            out = (self.host_out if self.host_io else self.dev_out)[slot]
            mask = (self.host_mask if self.host_io else self.dev_mask)[slot].view(torch.bool) if self.want_mask else None
            v = self._views[slot] = (list(out[: 9 * B].view(B, 3, 3).unbind(0)), mask, out[10 * B: 11 * B])
        return v

    def drain(self):
        for s in range(self.slots):
            if self.busy[s]:
                self.done[s].synchronize()
                self.busy[s] = False

    def after(self, event):
        """Order every stream of the service behind `event` (for timing from another stream)."""
        self.copy_stream.wait_event(event)
        for ks in self.compute:
            ks.wait_event(event)

    def join(self, stream):
        """Make `stream` wait for everything enqueued so far."""
        stream.wait_stream(self.copy_stream)
        for ks in self.compute:
            stream.wait_stream(ks)
