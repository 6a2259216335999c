"""autograd.Function wrappers giving the per-sample estimator plugins (the reference's
`estimate_model(matches[K,s,D])` call shape) gradients w.r.t. the minimal samples."""
from __future__ import annotations

import torch

from .. import ops


class E5AllSlots(torch.autograd.Function):
    """pts [K,5,4] -> models [K,10,3,3], nsol [K].  Backward: implicit-function adjoint per slot."""

    @staticmethod
    def forward(ctx, pts):
        models, nsol = ops.solve_e5(pts)
        ctx.save_for_backward(ops._f32(pts), models, nsol)
        ctx.mark_non_differentiable(nsol)
        return models[0], nsol[0]

    @staticmethod
    def backward(ctx, g_models, _):
        pts, models, nsol = ctx.saved_tensors
        g = torch.zeros_like(pts)
        gm = g_models.reshape(1, -1, ops.E5_SLOTS, 9)
        for s in range(ops.E5_SLOTS):
            sel = torch.where(nsol > s, torch.full_like(nsol, s), torch.full_like(nsol, -1))
            if (sel >= 0).any():
                g = g + ops.solve_e5_backward(pts, None, models, sel, gm[:, :, s].contiguous())[0]
        return g


class F8Solve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts):
        models, valid = ops.solve_f8(pts)
        ctx.save_for_backward(ops._f32(pts))
        ctx.mark_non_differentiable(valid)
        return models[0], valid[0]

    @staticmethod
    def backward(ctx, g_models, _):
        (pts,) = ctx.saved_tensors
        return ops.solve_f8_backward(pts, None, g_models.reshape(1, -1, 9).contiguous())[0]


class Rigid3Solve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, flag):
        models, valid = ops.solve_rigid3(pts, None, flag)
        ctx.save_for_backward(ops._f32(pts))
        ctx.flag = bool(flag)
        ctx.mark_non_differentiable(valid)
        return models[0], valid[0]

    @staticmethod
    def backward(ctx, g_models, _):
        (pts,) = ctx.saved_tensors
        return ops.solve_rigid3_backward(pts, None, g_models.reshape(1, -1, 16).contiguous(), ctx.flag)[0], None
