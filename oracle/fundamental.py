"""Oracle: normalised 8-point fundamental-matrix solver, and a CORRECT 7-point.

8-point restates `estimators/fundamental_matrix_estimator.py:177-260`
(`FundamentalMatrixEstimatorNew.normalize` + `estimate_non_minimal_model`).
The reference's 7-point (`:262-308`) is broken as shipped (F2 built
from null-space column 0, companion overwrites its own sub-diagonal), so the
7-point here follows the textbook algorithm the reference's comments describe
and is pinned by its own algebra (x2^T F x1 = 0 on the sample, det F = 0) --
"parity unpinned" against the reference for that one function.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import torch


def hartley_normalize(matches: torch.Tensor):
    """fundamental_matrix_estimator.py:177-217.  matches [K,s,4] ->
    (normalised [K,s,4], T1 [K,3,3], T2t [K,3,3]) where T2t is ALREADY the
    transpose of the second image's normalising transform (its translation is
    written into the last row, :213-214)."""
    K = matches.shape[0]
    mass = matches.mean(dim=1)
    c = matches - mass.unsqueeze(1)
    d1 = torch.linalg.norm(c[:, :, :2], dim=2).mean(dim=1)
    d2 = torch.linalg.norm(c[:, :, 2:], dim=2).mean(dim=1)
    r1 = math.sqrt(2) / d1
    r2 = math.sqrt(2) / d2
    n1 = c[:, :, :2] * r1.view(-1, 1, 1)
    n2 = c[:, :, 2:] * r2.view(-1, 1, 1)
    T1 = torch.zeros(K, 3, 3, dtype=matches.dtype)
    T2t = torch.zeros(K, 3, 3, dtype=matches.dtype)
    T1[:, 0, 0] = T1[:, 1, 1] = r1
    T2t[:, 0, 0] = T2t[:, 1, 1] = r2
    T1[:, 2, 2] = T2t[:, 2, 2] = 1
    T1[:, 0, 2] = -r1 * mass[:, 0]
    T1[:, 1, 2] = -r1 * mass[:, 1]
    T2t[:, 2, 0] = -r2 * mass[:, 2]
    T2t[:, 2, 1] = -r2 * mass[:, 3]
    return torch.cat((n1, n2), dim=2), T1, T2t


def f_rows(pts: torch.Tensor, weights=None) -> torch.Tensor:
    """Row (x1x2, x2y1, x2, y2x1, y2y1, y2, x1, y1, 1)  (fundamental_matrix_estimator.py:240-245):
    row-major vec of F with x2^T F x1 = 0."""
    x1, y1, x2, y2 = pts[..., 0], pts[..., 1], pts[..., 2], pts[..., 3]
    A = torch.stack((x1 * x2, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, torch.ones_like(x1)), dim=-1)
    if weights is not None:
        A = weights.unsqueeze(-1) * A
    return A


def eight_point(matches: torch.Tensor, weights=None) -> torch.Tensor:
    """matches [K,s>=8,4] -> F [K,3,3]: smallest right singular vector of A^T A on
    normalised points, de-normalised F = T2^T Fn T1 in a K-trip loop
    (fundamental_matrix_estimator.py:230-260).  No rank-2 projection, no rescale."""
    norm, T1, T2t = hartley_normalize(matches)
    A = f_rows(norm, weights)
    _, _, vh = torch.linalg.svd(A.transpose(-1, -2) @ A)
    F = vh[:, -1, :].reshape(-1, 3, 3).clone()
    for i in range(F.shape[0]):                                   # :256-258 python loop over K
        F[i] = torch.mm(T2t[i], torch.mm(F[i].clone(), T1[i]))
    return F


def seven_point(matches: torch.Tensor):
    """Correct 7-point: F = a F1 + (1 - a) F2 over the 2-dim null space, det F = 0
    cubic in a, up to three real roots.  matches [K,7,4] -> (F [K,3,3,3],
    valid [K,3] bool); invalid slots hold the identity (as the reference pads,
    fundamental_matrix_estimator.py:304-307).  Each F is scaled to unit norm.
    """
    K = matches.shape[0]
    A = f_rows(matches).to(torch.float64)
    _, _, vh = torch.linalg.svd(A, full_matrices=True)
    F1 = vh[:, -1, :].reshape(K, 3, 3)
    F2 = vh[:, -2, :].reshape(K, 3, 3)

    def det_at(a):
        return torch.linalg.det(a * F1 + (1 - a) * F2)

    # interpolate the cubic c0 + c1 a + c2 a^2 + c3 a^3 from 4 samples (the
    # scheme sketched at fundamental_matrix_estimator.py:219-227)
    d0, d1, dm1, d2, dm2 = det_at(0.0), det_at(1.0), det_at(-1.0), det_at(2.0), det_at(-2.0)
    c0 = d0
    c2 = 0.5 * (d1 + dm1) - d0
    c3 = ((d2 - dm2) / 2 - (d1 - dm1)) / 6
    c1 = (d1 - dm1) / 2 - c3
    out = torch.eye(3, dtype=torch.float64).repeat(K, 3, 1, 1)
    valid = torch.zeros(K, 3, dtype=torch.bool)
    for k in range(K):
        comp = torch.zeros(3, 3, dtype=torch.float64)
        comp[1, 0] = comp[2, 1] = 1.0
        comp[:, 2] = -torch.stack((c0[k], c1[k], c2[k])) / c3[k]
        ev = torch.linalg.eigvals(comp)
        real = ev[ev.imag.abs() < 1e-9 * (1 + ev.real.abs())].real.sort().values
        for s, a in enumerate(real):
            Fk = a * F1[k] + (1 - a) * F2[k]
            out[k, s] = Fk / Fk.norm()
            valid[k, s] = True
    return out.to(matches.dtype), valid
