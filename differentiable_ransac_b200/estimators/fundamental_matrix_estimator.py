"""FundamentalMatrixEstimatorNew -- host mirror of `estimators/fundamental_matrix_estimator.py:161-308`."""
from __future__ import annotations

import torch

from .. import ops
from ._autograd import F8Solve


def normalized_eight_point_torch(matches, weights=None):
    """Non-minimal (n > 8) normalised 8-point with torch ops on the device -- the final refit of
    `ransac.py:150-155`, once per pair and outside the hot loop.  matches [1,n,4] -> [1,3,3]."""
    m = matches.float()
    mass = m.mean(1, keepdim=True)
    c = m - mass
    r1 = (2 ** 0.5) / c[..., :2].norm(dim=-1).mean(1)
    r2 = (2 ** 0.5) / c[..., 2:].norm(dim=-1).mean(1)
    x1, y1 = c[..., 0] * r1[:, None], c[..., 1] * r1[:, None]
    x2, y2 = c[..., 2] * r2[:, None], c[..., 3] * r2[:, None]
    A = torch.stack((x1 * x2, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, torch.ones_like(x1)), -1)
    if weights is not None:
        A = weights.reshape(1, -1, 1).float() * A
    _, vecs = torch.linalg.eigh(A.transpose(-1, -2) @ A)
    Fn = vecs[..., 0].reshape(-1, 3, 3)
    B = m.shape[0]
    T1 = torch.zeros(B, 3, 3, device=m.device)
    T2 = torch.zeros(B, 3, 3, device=m.device)
    T1[:, 0, 0] = T1[:, 1, 1] = r1
    T2[:, 0, 0] = T2[:, 1, 1] = r2
    T1[:, 2, 2] = T2[:, 2, 2] = 1
    T1[:, 0, 2], T1[:, 1, 2] = -r1 * mass[:, 0, 0], -r1 * mass[:, 0, 1]
    T2[:, 0, 2], T2[:, 1, 2] = -r2 * mass[:, 0, 2], -r2 * mass[:, 0, 3]
    return T2.transpose(-1, -2) @ Fn @ T1


class FundamentalMatrixEstimatorNew:
    def __init__(self, device="cuda", weighted=0):
        self.sample_size = 7
        self.device = device
        self.weighted = weighted
        self.eps = 1e-8
        self.last_nsol = None

    def estimate_model(self, matches, weights=None):
        s = matches.shape[1]
        if s == 8:                                    # normalised 8-point (:230-260)
            models, _ = F8Solve.apply(matches.float())
            return models.to(matches.dtype)
        if s == 7:                                    # correct 7-point; the reference's is broken
            models, nsol = ops.solve_f7(matches.float())
            self.last_nsol = nsol[0]
            return models[0].reshape(-1, 3, 3).to(matches.dtype)
        if s > 8:
            return normalized_eight_point_torch(matches, weights).to(matches.dtype)
        return None
