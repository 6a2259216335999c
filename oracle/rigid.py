"""Oracle: 3-point rigid-transformation solver (both `flag` branches).

Restates `estimators/rigid_transformation_SVD_based_solver.py:11-74`.  With the
default flag=True the reference takes the SVD of cov^T cov (symmetric PSD), so
R = V U^T collapses to the identity and the "solver" reduces to the centroid
translation; flag=False is the Kabsch-style branch (it returns the
transpose of the usual Kabsch rotation).  Both are restated verbatim because
`RANSAC3D` calls the default (`ransac.py:367`).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import torch


def estimate(points: torch.Tensor, flag: bool = True):
    """points [K,n>=3,6] = [P | Q] -> (model [K,4,4], R [K,3,3], t [K,3], scale [K])."""
    assert points.shape[-1] == 6 and points.shape[-2] >= 3
    n = points.shape[1]
    centroid = points.mean(dim=1)
    c = points - centroid[:, None, :]
    avg0 = c[:, :, 0:3].pow(2).sum(-1).sqrt().sum(-1) / n
    avg1 = c[:, :, 3:6].pow(2).sum(-1).sqrt().sum(-1) / n
    sqrt3 = torch.sqrt(torch.tensor(3.0))                      # float32 constant, as :9
    c0 = c.transpose(-1, -2)[:, 0:3, :] * (sqrt3 / avg0)[:, None, None]
    c1 = c.transpose(-1, -2)[:, 3:6, :] * (sqrt3 / avg1)[:, None, None]
    cov = c0 @ c1.transpose(-1, -2)
    ok = ~torch.isnan(cov).flatten(1).any(dim=1)                # :45 nan_filter
    if flag:
        u, _, vh = torch.linalg.svd(cov.transpose(-1, -2) @ cov)
    else:
        u, _, vh = torch.linalg.svd(cov.transpose(-1, -2))
    v = vh.clone().transpose(-1, -2)
    R = v @ u.transpose(-1, -2)
    neg = torch.linalg.det(R) < 0
    if neg.any():                                               # :59-62 reflection fix on V's last column
        v[neg, :, 2] = -v[neg, :, 2]
        R = v @ u.transpose(-1, -2)
    scale = avg1 / avg0
    # :66 -- broadcasts -centroid over ROWS and sums over rows: t_j = -c0_j * sum_i R_ij + c1_j
    t = torch.sum(R * (-centroid[:, None, 0:3]), dim=1) + centroid[:, 3:6]
    bottom = torch.tensor([[0, 0, 0, 1]], dtype=R.dtype).repeat(R.shape[0], 1, 1)
    model = torch.cat((torch.cat((R, t.unsqueeze(-1)), dim=-1), bottom), dim=1)
    return model[ok], R[ok], t[ok], scale[ok]
