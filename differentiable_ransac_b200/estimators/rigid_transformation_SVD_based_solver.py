"""RigidTransformationSVDBasedSolver -- host mirror of
`estimators/rigid_transformation_SVD_based_solver.py:4-89` over the CUDA kernels."""
from __future__ import annotations

import torch

from .. import engine
from ._autograd import Rigid3Solve


class RigidTransformationSVDBasedSolver:
    def __init__(self, data_type=torch.float32, device="cuda"):
        self.data_type = data_type
        self.device = device
        self.sample_size = 3

    def estimate_model(self, data, weights=None, sample_indices=None, flag=True):
        """data [K,3,6] -> (model [K',4,4], R, t, scale) with invalid samples dropped (:45,74)."""
        assert data.shape[-1] == 6 and data.shape[-2] == 3, "the CUDA solver is the minimal 3-point one"
        if sample_indices is not None:
            data = torch.index_select(data, 0, sample_indices)
        model, valid = Rigid3Solve.apply(data.float(), bool(flag))
        keep = valid.bool()
        model = model[keep]
        c = data[keep].float()
        c = c - c.mean(1, keepdim=True)
        scale = c[..., 3:].norm(dim=-1).mean(1) / c[..., :3].norm(dim=-1).mean(1)
        return model, model[:, :3, :3], model[:, :3, 3], scale

    def squared_residual(self, pts1, pts2, descriptor, threshold=0.03):
        """pts1, pts2 [N,3]; descriptor [K,4,3] = model[:, :3, :]^T -> (sum d2 [K], mean d2, LazyCount).
        The [K,N] inlier mask of the reference is replaced by per-model inlier counts."""
        K = descriptor.shape[0]
        models = torch.zeros(1, K, 4, 4, device=descriptor.device)
        models[0, :, :3, :] = descriptor.transpose(-1, -2)
        models[0, :, 3, 3] = 1.0
        points = torch.cat((pts1, pts2), -1)[None].float()
        res = engine.RigidResidual.apply(points, models)[0]
        return res, res.sum() / (K * pts1.shape[0]), None
