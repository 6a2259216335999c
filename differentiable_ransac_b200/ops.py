"""Thin torch wrappers over the C ABI (include/drb.h): one function per entry point.
Tensors are CUDA fp32 / int32, contiguous; every call runs on torch's current stream.
PyTorch here is plumbing (device memory + streams) -- all arithmetic is in libdrb.so."""
from __future__ import annotations

import ctypes
import os

import torch

import threading

from . import _lib

E5_SLOTS = 10
F7_SLOTS = 3

# Every tensor whose address goes into a launch is kept alive until that launch has been enqueued: a converted
# copy made inline (`_p(_f32(x))`) would otherwise be freed as soon as its pointer is taken, and the caching
# allocator could hand the same block to the NEXT inline copy of the same call.  After the launch the stream
# orders any reuse behind the kernel, so `check` drops the references.
_ARGS = threading.local()


def check(status: int, what: str):
    getattr(_ARGS, "alive", []).clear()
    _lib.check(status, what)


def _p(t):
    if t is None:
        return None
    if not hasattr(_ARGS, "alive"):
        _ARGS.alive = []
    _ARGS.alive.append(t)
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.device.type != "cuda":
        raise _lib.DrbError("libdrb operates on CUDA tensors only (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _i32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


# ---- sampler -----------------------------------------------------------------------------------
def _offset_dev(offset_dev, device):
    if offset_dev is not None and (offset_dev.dtype != torch.int64 or offset_dev.device != device):
        raise _lib.DrbError("offset_dev must be an int64 tensor on the device of `logits`")
    return offset_dev


def sample(logits, K, s, tau=1.0, noise=None, seed=0, offset=0, want_lse=False, want_noise=False, offset_dev=None):
    """logits [B,N] -> idx [B,K,s] int32 (ascending per row), lse [B,K] | None,
    sel_key [B,K,s] | None, noise_out [B,K,N] | None.
    offset_dev: optional int64 [1] CUDA tensor added to `offset` on the device (CUDA-graph replay)."""
    logits = _f32(logits)
    B, N = logits.shape
    dev = logits.device
    noise = None if noise is None else _f32(noise)
    if noise is not None and tuple(noise.shape) != (B, K, N):
        raise ValueError(f"noise must be [B,K,N] = {(B, K, N)}, got {tuple(noise.shape)}")
    idx = torch.empty(B, K, s, dtype=torch.int32, device=dev)
    lse = torch.empty(B, K, dtype=torch.float32, device=dev) if want_lse else None
    sel_key = torch.empty(B, K, s, dtype=torch.float32, device=dev) if want_lse else None
    noise_out = torch.empty(B, K, N, dtype=torch.float32, device=dev) if want_noise else None
    lib = _lib.load()
    check(lib.drb_sample(_p(logits), _p(noise), seed, offset, _p(_offset_dev(offset_dev, dev)), float(tau), B, K, N, s,
                         _p(idx), _p(lse), _p(sel_key), _p(noise_out), _stream()), "drb_sample")
    return idx, lse, sel_key, noise_out


def sample_sets(logits, K, s, seed=0, offset=0, offset_dev=None):
    """Test-mode set sampler (Plackett-Luce by inverse CDF, no Gumbel noise): logits [B,N] -> idx [B,K,s].
    Falls back to the Gumbel-race kernel when one pair's prefix sums do not fit in shared memory.
    offset_dev: optional int64 [1] CUDA tensor added to `offset` on the device (for CUDA-graph replay)."""
    logits = _f32(logits)
    B, N = logits.shape
    idx = torch.empty(B, K, s, dtype=torch.int32, device=logits.device)
    lib = _lib.load()
    _offset_dev(offset_dev, logits.device)
    rc = lib.drb_sample_sets(_p(logits), seed, offset, _p(offset_dev), B, K, N, s, _p(idx), _stream())
    if rc == -3 and s in (3, 5, 7, 8) and offset_dev is None:
        return sample(logits, K, s, 1.0, None, seed, offset)[0]
    check(rc, "drb_sample_sets")
    return idx


def sample_backward(logits, idx, lse, sel_key, g_sel, tau=1.0, noise=None, seed=0, offset=0, offset_dev=None):
    logits = _f32(logits)
    B, N = logits.shape
    _, K, s = idx.shape
    noise = None if noise is None else _f32(noise)
    g_sel = _f32(g_sel)
    grad = torch.zeros(B, N, dtype=torch.float32, device=logits.device)
    scratch = torch.empty(B, K, dtype=torch.float32, device=logits.device)
    lib = _lib.load()
    check(lib.drb_sample_backward(_p(logits), _p(noise), seed, offset, _p(_offset_dev(offset_dev, logits.device)),
                                  float(tau), B, K, N, s, _p(idx), _p(lse),
                                  _p(sel_key), _p(g_sel), _p(scratch), _p(grad), _stream()), "drb_sample_backward")
    return grad


def gather_backward(matches, idx, g_pts, want_grad_matches=True):
    """g_pts [B,K,s,D] -> g_sel [B,K,s], grad_matches [B,N,D] | None."""
    matches = _f32(matches)
    B, N, D = matches.shape
    _, K, s = idx.shape
    g_pts = _f32(g_pts)
    g_sel = torch.empty(B, K, s, dtype=torch.float32, device=matches.device)
    gm = torch.zeros_like(matches) if want_grad_matches else None
    lib = _lib.load()
    check(lib.drb_gather_backward(_p(matches), _p(idx), _p(g_pts), B, K, N, s, D, _p(g_sel), _p(gm), _stream()),
          "drb_gather_backward")
    return g_sel, gm


# ---- solvers -----------------------------------------------------------------------------------
def _rows(matches, idx, s, D):
    """Resolve the (gathered | indexed) calling convention -> (matches, idx, B, K, N)."""
    matches = _f32(matches)
    if idx is None:
        if matches.dim() != 3 or matches.shape[1] != s or matches.shape[2] != D:
            raise ValueError(f"gathered minimal samples must be [M,{s},{D}], got {tuple(matches.shape)}")
        return matches, None, 1, matches.shape[0], 0
    idx = _i32(idx)
    B, N, _ = matches.shape
    if idx.shape[0] != B or idx.shape[2] != s:
        raise ValueError("idx must be [B,K,s]")
    return matches, idx, B, idx.shape[1], N


def zeroed_counters(B, device):
    """One fill for both zero-initialised scratch arrays of the test pipeline:
    best_packed [B] (u64 as int64) and ccount [B] int32."""
    z = torch.zeros(2 * B, dtype=torch.int64, device=device)
    return z[:B], z[B:].view(torch.int32)[:B]


_WS_CACHE = {}


def score_workspace(B, M, N, device):
    """Workspace of drb_score_msac_stream for (B, M, N): an int64 tensor whose leading queue / arrival counters
    are zero.  The kernel leaves them zero, and launches on one stream are ordered, so the buffer is cached
    per (device, stream, size) -- two streams never share one."""
    lib = _lib.load()
    nbytes = int(lib.drb_score_msac_workspace_bytes(int(B), int(M), int(N)))
    zbytes = int(lib.drb_score_msac_workspace_zeroed_bytes(int(B), int(M)))
    dev = torch.device(device)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, nbytes, zbytes,
           torch.cuda.is_current_stream_capturing())
    ws = _WS_CACHE.get(key)
    if ws is None:
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
        ws[: zbytes // 8].zero_()
        if len(_WS_CACHE) > 64:
            _WS_CACHE.clear()
        _WS_CACHE[key] = ws
    return ws


def solve_e5(matches, idx=None, compact=False, ccount=None):
    """-> models [B,K,10,3,3], nsol [B,K] int32 (, cmodels [B,K*10,9], cids [B,K*10], ccount [B])."""
    matches, idx, B, K, N = _rows(matches, idx, 5, 4)
    dev = matches.device
    models = torch.empty(B, K, E5_SLOTS, 3, 3, dtype=torch.float32, device=dev)
    nsol = torch.empty(B, K, dtype=torch.int32, device=dev)
    cm = cid = cc = None
    if compact:
        cm = torch.empty(B, K * E5_SLOTS, 9, dtype=torch.float32, device=dev)
        cid = torch.empty(B, K * E5_SLOTS, dtype=torch.int32, device=dev)
        cc = ccount if ccount is not None else torch.zeros(B, dtype=torch.int32, device=dev)
    lib = _lib.load()
    check(lib.drb_solve_e5(_p(matches), _p(idx), B, K, N, _p(models), _p(nsol), _p(cm), _p(cid), _p(cc), _stream()),
          "drb_solve_e5")
    if compact:
        return models, nsol, cm, cid, cc
    return models, nsol


def solve_e5_backward(matches, idx, models, sel, g_model):
    matches, idx, B, K, N = _rows(matches, idx, 5, 4)
    g_pts = torch.empty(B, K, 5, 4, dtype=torch.float32, device=matches.device)
    lib = _lib.load()
    check(lib.drb_solve_e5_backward(_p(matches), _p(idx), B, K, N, _p(_f32(models)), _p(_i32(sel)), _p(_f32(g_model)),
                                    _p(g_pts), _stream()), "drb_solve_e5_backward")
    return g_pts


def solve_e5_select(matches, idx, gt, sign_invariant=True, want_models=False):
    """Train mode in one launch: five-point solve + the slot closest to gt [B,3,3].
    -> sel [B,K] int32 (-1: none), chosen [B,K,3,3], nsol [B,K] (, models [B,K,10,3,3] when want_models)."""
    matches, idx, B, K, N = _rows(matches, idx, 5, 4)
    dev = matches.device
    models = torch.empty(B, K, E5_SLOTS, 3, 3, dtype=torch.float32, device=dev) if want_models else None
    nsol = torch.empty(B, K, dtype=torch.int32, device=dev)
    sel = torch.empty(B, K, dtype=torch.int32, device=dev)
    chosen = torch.empty(B, K, 3, 3, dtype=torch.float32, device=dev)
    lib = _lib.load()
    check(lib.drb_solve_e5_select(_p(matches), _p(idx), _p(_f32(gt).reshape(B, 9)), int(bool(sign_invariant)), B, K, N,
                                  _p(models), _p(nsol), _p(sel), _p(chosen), _stream()), "drb_solve_e5_select")
    return sel, chosen, nsol, models


def solve_e5_backward_chosen(matches, idx, chosen, sel, g_model):
    """`solve_e5_backward` from the chosen models [B,K,3,3] alone."""
    matches, idx, B, K, N = _rows(matches, idx, 5, 4)
    g_pts = torch.empty(B, K, 5, 4, dtype=torch.float32, device=matches.device)
    lib = _lib.load()
    check(lib.drb_solve_e5_backward_chosen(_p(matches), _p(idx), B, K, N, _p(_f32(chosen)), _p(_i32(sel)),
                                           _p(_f32(g_model)), _p(g_pts), _stream()), "drb_solve_e5_backward_chosen")
    return g_pts


def select_closest(models, nsol, gt, sign_invariant=True):
    """models [B,K,S,3,3], nsol [B,K], gt [B,3,3] -> sel [B,K] int32, chosen [B,K,3,3]."""
    models = _f32(models)
    B, K, S = models.shape[:3]
    sel = torch.empty(B, K, dtype=torch.int32, device=models.device)
    chosen = torch.empty(B, K, 3, 3, dtype=torch.float32, device=models.device)
    lib = _lib.load()
    check(lib.drb_select_closest(_p(models), _p(_i32(nsol)), _p(_f32(gt)), B, K, S, int(bool(sign_invariant)),
                                 _p(sel), _p(chosen), _stream()), "drb_select_closest")
    return sel, chosen


def solve_f8(matches, idx=None):
    matches, idx, B, K, N = _rows(matches, idx, 8, 4)
    models = torch.empty(B, K, 3, 3, dtype=torch.float32, device=matches.device)
    valid = torch.empty(B, K, dtype=torch.uint8, device=matches.device)
    lib = _lib.load()
    check(lib.drb_solve_f8(_p(matches), _p(idx), B, K, N, _p(models), _p(valid), _stream()), "drb_solve_f8")
    return models, valid


def solve_f8_backward(matches, idx, g_model, models=None):
    """`models` (the forward's output, optional) fixes the sign of the null vector the backward recomputes;
    without it the kernel re-runs the forward for that."""
    matches, idx, B, K, N = _rows(matches, idx, 8, 4)
    g_pts = torch.empty(B, K, 8, 4, dtype=torch.float32, device=matches.device)
    lib = _lib.load()
    check(lib.drb_solve_f8_backward(_p(matches), _p(idx), B, K, N, _p(None if models is None else _f32(models)),
                                    _p(_f32(g_model)), _p(g_pts), _stream()), "drb_solve_f8_backward")
    return g_pts


def solve_f7(matches, idx=None):
    matches, idx, B, K, N = _rows(matches, idx, 7, 4)
    models = torch.empty(B, K, F7_SLOTS, 3, 3, dtype=torch.float32, device=matches.device)
    nsol = torch.empty(B, K, dtype=torch.int32, device=matches.device)
    lib = _lib.load()
    check(lib.drb_solve_f7(_p(matches), _p(idx), B, K, N, _p(models), _p(nsol), _stream()), "drb_solve_f7")
    return models, nsol


def _refit(name, slots, matches, mask, weights):
    matches = _f32(matches)
    B, N, _ = matches.shape
    models = torch.empty(B, slots, 3, 3, dtype=torch.float32, device=matches.device)
    nsol = torch.empty(B, dtype=torch.int32, device=matches.device)
    mk = None if mask is None else mask.to(torch.uint8).contiguous()
    w = None if weights is None else _f32(weights)
    lib = _lib.load()
    check(getattr(lib, name)(_p(matches), _p(mk), _p(w), B, N, _p(models), _p(nsol), _stream()), name)
    return models, nsol


def refit_e5(matches, mask=None, weights=None):
    """Essential matrices of the selected correspondences of each pair (non-minimal five-point):
    matches [B,N,4], mask [B,N] or None -> models [B,10,3,3] (identity beyond nsol), nsol [B]."""
    return _refit("drb_refit_e5", E5_SLOTS, matches, mask, weights)


def refit_f8(matches, mask=None, weights=None):
    """Hartley-normalised eight-point on the selected correspondences: -> F [B,1,3,3], ok [B]."""
    return _refit("drb_refit_f8", 1, matches, mask, weights)


def recover_pose(E, matches, npts=None, R_gt=None, t_gt=None, dist=50.0, want_mask=True):
    """E [B,M,3,3] (or [B,3,3]), matches [B,N,4] normalised coordinates -> dict(R [B,M,3,3], t [B,M,3],
    mask [B,M,N] bool | None, ngood [B,M], err [B,M,2] degrees | None)."""
    matches = _f32(matches)
    B, N, _ = matches.shape
    E = _f32(E).reshape(B, -1, 9)
    M = E.shape[1]
    dev = matches.device
    R = torch.empty(B, M, 3, 3, dtype=torch.float32, device=dev)
    t = torch.empty(B, M, 3, dtype=torch.float32, device=dev)
    mask = torch.empty(B, M, N, dtype=torch.uint8, device=dev) if want_mask else None
    ngood = torch.empty(B, M, dtype=torch.int32, device=dev)
    have_gt = R_gt is not None and t_gt is not None
    err = torch.empty(B, M, 2, dtype=torch.float32, device=dev) if have_gt else None
    votes = torch.empty(B, M, 4, dtype=torch.int32, device=dev)
    lib = _lib.load()
    check(lib.drb_recover_pose(_p(E), _p(matches), _p(None if npts is None else _i32(npts)),
                               _p(_f32(R_gt).reshape(B, 9) if have_gt else None),
                               _p(_f32(t_gt).reshape(B, 3) if have_gt else None), B, M, N, float(dist), _p(votes), _p(R),
                               _p(t), _p(mask), _p(ngood), _p(err), _stream()), "drb_recover_pose")
    return dict(R=R, t=t, mask=None if mask is None else mask.view(torch.bool), ngood=ngood, err=err, votes=votes)


def pose_loss(E, matches, R_gt, t_gt, npts=None, dist=50.0, want_grad=True):
    """E [B,M,3,3], matches [B,N,4], R_gt [B,3,3], t_gt [B,3] -> err [B,M,2] degrees (Horn decomposition, as
    PoseLoss), grad [B,M,3,3] = d((err_R + err_t)/2)/dE or None."""
    matches = _f32(matches)
    B, N, _ = matches.shape
    E = _f32(E).reshape(B, -1, 9)
    M = E.shape[1]
    err = torch.empty(B, M, 2, dtype=torch.float32, device=matches.device)
    grad = torch.empty(B, M, 3, 3, dtype=torch.float32, device=matches.device) if want_grad else None
    votes = torch.empty(B, M, 4, dtype=torch.int32, device=matches.device)
    lib = _lib.load()
    check(lib.drb_pose_loss(_p(E), _p(matches), _p(None if npts is None else _i32(npts)), _p(_f32(R_gt).reshape(B, 9)),
                            _p(_f32(t_gt).reshape(B, 3)), B, M, N, float(dist), _p(votes), _p(err), _p(grad),
                            _stream()), "drb_pose_loss")
    return err, grad


def solve_rigid3(points, idx=None, flag=True):
    points, idx, B, K, N = _rows(points, idx, 3, 6)
    models = torch.empty(B, K, 4, 4, dtype=torch.float32, device=points.device)
    valid = torch.empty(B, K, dtype=torch.uint8, device=points.device)
    lib = _lib.load()
    check(lib.drb_solve_rigid3(_p(points), _p(idx), B, K, N, int(bool(flag)), _p(models), _p(valid), _stream()),
          "drb_solve_rigid3")
    return models, valid


def solve_rigid3_backward(points, idx, g_model, flag=True):
    points, idx, B, K, N = _rows(points, idx, 3, 6)
    g_pts = torch.empty(B, K, 3, 6, dtype=torch.float32, device=points.device)
    lib = _lib.load()
    check(lib.drb_solve_rigid3_backward(_p(points), _p(idx), B, K, N, int(bool(flag)), None, _p(_f32(g_model)),
                                        _p(g_pts), _stream()), "drb_solve_rigid3_backward")
    return g_pts


# ---- scoring -----------------------------------------------------------------------------------
_MSAC_KERNEL = os.environ.get("DRB_MSAC_KERNEL", "stream")
# tensor-core scorer variants -> the `words` argument of drb_score_msac_tc: operand split (2 TF32 / 3 BF16 words),
# "p" = one reciprocal per model pair (+16), "_e16" = 16 epilogue warps instead of 8 (+32)
# a trailing "p": one reciprocal per pair of neighbouring models (+16); "q": the same with the threshold folded into
# the denominator rows and the clamp moved to the ALU pipe (+128, msac_tc_layout.cuh::model_rows_folded)
_TC_WORDS = {"tc": 3, "tc_bf16": 3, "tc_tf32": 2, "tc_bf16p": 3 + 16, "tc_tf32p": 2 + 16, "tc_bf16q": 3 + 16 + 128,
             "tc_tf32q": 2 + 16 + 128}
_TC_WORDS.update({k + "_e16": v + 32 for k, v in list(_TC_WORDS.items()) if k != "tc"})
# "*_s": the slim build of the pair variant (+256: 128 registers, three stages; leaves room for a five-point CTA on the SM)
_TC_WORDS.update({"tc_bf16p_s": 3 + 16 + 256, "tc_tf32p_s": 2 + 16 + 256})
# "*2": two SMs per tile (+512: tcgen05 cta_group::2, csrc/score_tc_pair.cu)
_TC_WORDS.update({"tc_bf16p2": 3 + 16 + 512, "tc_tf32p2": 2 + 16 + 512})
# "tc2_*": the model-stationary arrangement (csrc/score_tc2.cu, +64)
_TC_WORDS.update({"tc2_tf32": 2 + 64, "tc2_bf16": 3 + 64, "tc2_tf32_e16": 2 + 64 + 32, "tc2_bf16_e16": 3 + 64 + 32,
                  "tc2_tf32p": 2 + 64 + 16, "tc2_bf16p": 3 + 64 + 16, "tc2_tf32p_e16": 2 + 64 + 16 + 32,
                  "tc2_bf16p_e16": 3 + 64 + 16 + 32})


def score_msac(matches, models, thr, count=None, ids=None, want_scores=True, best=None, kernel=None):
    """matches [B,N,4], models [B,M,9|3,3], thr [B] -> scores [B,M] | None, best_packed [B] (int64 view of u64).
    `best` may be a caller-zeroed [B] int64 buffer.  kernel="stream": the evenly split persistent grid
    (drb_score_msac_stream); "block": one CTA per (pair, 32 models) (drb_score_msac); "tc_*": the tensor-core scorer
    (drb_score_msac_tc, _TC_WORDS above; the pipelined service's default, every variant pinned to the fp64 oracle)."""
    kernel = kernel or _MSAC_KERNEL
    matches = _f32(matches)
    B, N, _ = matches.shape
    models = _f32(models).reshape(B, -1, 9)
    M = models.shape[1]
    thr = _f32(thr).reshape(B)
    scores = torch.empty(B, M, dtype=torch.float32, device=matches.device) if want_scores else None
    if best is None:
        best = torch.zeros(B, dtype=torch.int64, device=matches.device)
    lib = _lib.load()
    count = None if count is None else _i32(count)
    ids = None if ids is None else _i32(ids)
    if kernel == "stream":
        ws = score_workspace(B, M, N, matches.device)
        check(lib.drb_score_msac_stream(_p(matches), _p(models), _p(count), _p(ids), _p(thr), B, M, N, _p(scores),
                                        _p(best), _p(ws), ws.numel() * 8, _stream()), "drb_score_msac_stream")
    elif kernel == "block":
        check(lib.drb_score_msac(_p(matches), _p(models), _p(count), _p(ids), _p(thr), B, M, N, _p(scores), _p(best),
                                 _stream()), "drb_score_msac")
    elif kernel in _TC_WORDS:
        # tensor-core scorer (csrc/score_tc.cu, score_tc2.cu), DESIGN.md section 10.
        # "tc" = "tc_bf16": three BF16 words per operand (fp32-level scores); "tc_tf32": two TF32 words;
        # a trailing "p": one reciprocal per model pair; "_e16": 16 epilogue warps; "_s": the slim build
        nbytes = int(lib.drb_score_msac_tc_workspace_bytes(B, N))
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=matches.device)
        check(lib.drb_score_msac_tc(_p(matches), _p(models), _p(count), _p(ids), _p(thr), B, M, N,
                                    _TC_WORDS[kernel], _p(scores), _p(best), _p(ws), ws.numel() * 8,
                                    _stream()), "drb_score_msac_tc")
    else:
        raise _lib.DrbError(f"unknown MSAC kernel {kernel!r}")
    return scores, best


def adaptive_select(matches, models_dense, scores, thr, span, rbs, max_iterations, sample_size, confidence=0.999,
                    eps=1e-5, count=None, ids=None):
    """Replay of the reference's chunked loop with adaptive exit on per-model scores (see drb.h):
    -> best_packed [B] (for best_finalize), iterations [B], chunk_best [B,C], chunk_ninl [B,C]."""
    matches = _f32(matches)
    B, N, _ = matches.shape
    md = _f32(models_dense).reshape(B, -1, 9)
    scores = _f32(scores).reshape(B, -1)
    C = -(-int(max_iterations) // int(rbs))
    dev = matches.device
    chunk_best = torch.empty(B, C, dtype=torch.int64, device=dev)
    chunk_ninl = torch.empty(B, C, dtype=torch.int32, device=dev)
    best = torch.empty(B, dtype=torch.int64, device=dev)
    its = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load()
    check(lib.drb_adaptive_select(_p(matches), _p(md), _p(scores), _p(None if count is None else _i32(count)),
                                  _p(None if ids is None else _i32(ids)), _p(_f32(thr).reshape(B)), B, scores.shape[1],
                                  md.shape[1], N, int(span), int(rbs), int(max_iterations), int(sample_size),
                                  float(confidence), float(eps), _p(chunk_best), _p(chunk_ninl), _p(best), _p(its),
                                  _stream()), "drb_adaptive_select")
    return best, its, chunk_best, chunk_ninl


def best_finalize(matches, models_dense, best_packed, thr, want_mask=True):
    matches = _f32(matches)
    B, N, _ = matches.shape
    md = _f32(models_dense).reshape(B, -1, 9)
    dev = matches.device
    best_id = torch.empty(B, dtype=torch.int32, device=dev)
    best_score = torch.empty(B, dtype=torch.float32, device=dev)
    best_model = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    mask = torch.empty(B, N, dtype=torch.uint8, device=dev) if want_mask else None
    ninl = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load()
    check(lib.drb_best_finalize(_p(matches), _p(md), _p(best_packed), _p(_f32(thr).reshape(B)), B, md.shape[1], N,
                                _p(best_id), _p(best_score), _p(best_model), _p(mask), _p(ninl), _stream()),
          "drb_best_finalize")
    return best_id, best_score, best_model, mask, ninl


def episym_forward(pts, models, npts=None, mvalid=None):
    """pts [B,P,4], models [B,K,3,3] -> row_sum [B,K] = sum_p min(episym, 1)."""
    pts = _f32(pts)
    B, P, _ = pts.shape
    models = _f32(models).reshape(B, -1, 9)
    K = models.shape[1]
    out = torch.empty(B, K, dtype=torch.float32, device=pts.device)
    lib = _lib.load()
    check(lib.drb_episym_forward(_p(pts), _p(None if npts is None else _i32(npts)), _p(models),
                                 _p(None if mvalid is None else mvalid.to(torch.uint8).contiguous()), B, K, P, _p(out),
                                 _stream()), "drb_episym_forward")
    return out


def episym_backward(pts, models, g_row, npts=None, mvalid=None):
    pts = _f32(pts)
    B, P, _ = pts.shape
    models = _f32(models).reshape(B, -1, 9)
    K = models.shape[1]
    out = torch.empty(B, K, 3, 3, dtype=torch.float32, device=pts.device)
    lib = _lib.load()
    check(lib.drb_episym_backward(_p(pts), _p(None if npts is None else _i32(npts)), _p(models),
                                  _p(None if mvalid is None else mvalid.to(torch.uint8).contiguous()),
                                  _p(_f32(g_row)), B, K, P, _p(out), _stream()), "drb_episym_backward")
    return out


def episym_forward_backward(pts, models, g_row, npts=None, mvalid=None):
    """One pass over the points: (row_sum [B,K], g_models [B,K,3,3]) = (episym_forward, episym_backward)."""
    pts = _f32(pts)
    B, P, _ = pts.shape
    models = _f32(models).reshape(B, -1, 9)
    K = models.shape[1]
    row = torch.empty(B, K, dtype=torch.float32, device=pts.device)
    out = torch.empty(B, K, 3, 3, dtype=torch.float32, device=pts.device)
    lib = _lib.load()
    check(lib.drb_episym_forward_backward(_p(pts), _p(None if npts is None else _i32(npts)), _p(models),
                                          _p(None if mvalid is None else mvalid.to(torch.uint8).contiguous()),
                                          _p(_f32(g_row)), B, K, P, _p(row), _p(out), _stream()),
          "drb_episym_forward_backward")
    return row, out


def rigid_residual_forward_backward(points, models, g_res):
    """One pass over the points: (res_sum [B,K], g_models [B,K,4,4])."""
    points = _f32(points)
    B, N, _ = points.shape
    models = _f32(models).reshape(B, -1, 16)
    K = models.shape[1]
    res = torch.empty(B, K, dtype=torch.float32, device=points.device)
    out = torch.empty(B, K, 4, 4, dtype=torch.float32, device=points.device)
    lib = _lib.load()
    check(lib.drb_rigid_residual_forward_backward(_p(points), _p(models), _p(_f32(g_res)), B, K, N, _p(res), _p(out),
                                                  _stream()), "drb_rigid_residual_forward_backward")
    return res, out


def rigid_residual_forward(points, models, threshold=0.03, want_ninl=True):
    points = _f32(points)
    B, N, _ = points.shape
    models = _f32(models).reshape(B, -1, 16)
    K = models.shape[1]
    res = torch.empty(B, K, dtype=torch.float32, device=points.device)
    ninl = torch.empty(B, K, dtype=torch.int32, device=points.device) if want_ninl else None
    lib = _lib.load()
    check(lib.drb_rigid_residual_forward(_p(points), _p(models), B, K, N, float(threshold), _p(res), _p(ninl),
                                         _stream()), "drb_rigid_residual_forward")
    return res, ninl


def rigid_residual_backward(points, models, g_res):
    points = _f32(points)
    B, N, _ = points.shape
    models = _f32(models).reshape(B, -1, 16)
    K = models.shape[1]
    out = torch.empty(B, K, 4, 4, dtype=torch.float32, device=points.device)
    lib = _lib.load()
    check(lib.drb_rigid_residual_backward(_p(points), _p(models), _p(_f32(g_res)), B, K, N, _p(out), _stream()),
          "drb_rigid_residual_backward")
    return out


# ---- the float64 chain (`-pr 2`; csrc/fp64_path.cu) -----------------------------------------------------------
def _f64(t):
    if t.device.type != "cuda":
        raise _lib.DrbError("the CUDA path needs CUDA tensors (there is no CPU fallback)")
    return t.detach().to(torch.float64).contiguous()


def solve_e5_f64(matches, idx):
    """matches [B,N,4] (any float dtype; computed in float64), idx [B,K,5] -> models [B,K,10,3,3] float64 (identity in
    the unused slots), nsol [B,K]."""
    m = _f64(matches)
    B, N, _ = m.shape
    idx = _i32(idx)
    K = idx.shape[1]
    models = torch.empty(B, K, E5_SLOTS, 3, 3, dtype=torch.float64, device=m.device)
    nsol = torch.empty(B, K, dtype=torch.int32, device=m.device)
    check(_lib.load().drb_solve_e5_f64(_p(m), _p(idx), B, K, N, _p(models), _p(nsol), _stream()), "drb_solve_e5_f64")
    return models, nsol


def score_msac_f64(matches, models, thr, nsol=None):
    """matches [B,N,4], models [B,K,S,3,3] (or [B,M,3,3] with nsol=None), thr [B] -> scores [B,K*S] float64 (-1 for a
    slot without a model)."""
    m = _f64(matches)
    B, N, _ = m.shape
    md = _f64(models)
    if nsol is not None:
        K, S = md.shape[1], md.shape[2]
    else:
        md = md.reshape(B, -1, 9)
        K, S = md.shape[1], 1
    scores = torch.empty(B, K * S, dtype=torch.float64, device=m.device)
    check(_lib.load().drb_score_msac_f64(_p(m), _p(md), _p(None if nsol is None else _i32(nsol)), _p(_f64(thr).reshape(B)),
                                         B, K, S, N, _p(scores), _stream()), "drb_score_msac_f64")
    return scores


def best_finalize_f64(matches, models, scores, thr):
    """-> best_id [B] (first maximum; -1 when no slot holds a model), best_score [B] f64, best_model [B,3,3] f64,
    mask [B,N] uint8, ninl [B]."""
    m = _f64(matches)
    B, N, _ = m.shape
    md = _f64(models).reshape(B, -1, 9)
    sc = _f64(scores).reshape(B, -1)
    M = md.shape[1]
    dev = m.device
    best_id = torch.empty(B, dtype=torch.int32, device=dev)
    best_score = torch.empty(B, dtype=torch.float64, device=dev)
    best_model = torch.empty(B, 3, 3, dtype=torch.float64, device=dev)
    mask = torch.empty(B, N, dtype=torch.uint8, device=dev)
    ninl = torch.empty(B, dtype=torch.int32, device=dev)
    check(_lib.load().drb_best_finalize_f64(_p(m), _p(md), _p(sc), _p(_f64(thr).reshape(B)), B, M, N, _p(best_id),
                                            _p(best_score), _p(best_model), _p(mask), _p(ninl), _stream()),
          "drb_best_finalize_f64")
    return best_id, best_score, best_model, mask, ninl
