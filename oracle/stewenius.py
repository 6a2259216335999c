"""Oracle: Stewenius 5-point essential-matrix solver (action-matrix / eig).

Restates `estimators/essential_matrix_estimator_stewenius.py:20-172`: SVD of the
5x9 epipolar matrix, the 10x20 constraint matrix in graded-reverse-lex monomial
order, `linalg.solve` elimination, the 10x10 action matrix and `linalg.eig`;
models are null_space @ Re(eigvec[-4:]) -- NOT normalised, and complex
eigenvectors contribute their real parts exactly as the reference does (:74-78).
The reference class cannot run as shipped; the golden fixtures
were produced with `estimator.device = 'cpu'` patched on.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import functools

import torch

from .nister import epipolar_rows

# graded reverse lexicographic orders used by stewenius.py:139-172
G1 = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
G2 = [(2, 0, 0), (1, 1, 0), (0, 2, 0), (1, 0, 1), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1),
      (0, 0, 0)]
G3 = [(3, 0, 0), (2, 1, 0), (1, 2, 0), (0, 3, 0), (2, 0, 1), (1, 1, 1), (0, 2, 1), (1, 0, 2), (0, 1, 2),
      (0, 0, 3), (2, 0, 0), (1, 1, 0), (0, 2, 0), (1, 0, 1), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0),
      (0, 0, 1), (0, 0, 0)]


@functools.lru_cache(maxsize=None)
def _structure(kind: str) -> torch.Tensor:
    left, right, out = {"12": (G1, G1, G2), "23": (G2, G1, G3)}[kind]
    T = torch.zeros(len(left), len(right), len(out), dtype=torch.float64)
    for ia, ma in enumerate(left):
        for ib, mb in enumerate(right):
            T[ia, ib, out.index(tuple(p + q for p, q in zip(ma, mb)))] = 1.0
    return T


def _m11(a, b):
    return torch.einsum("ka,kb,abc->kc", a, b, _structure("12").to(a.dtype))


def _m21(a, b):
    return torch.einsum("ka,kb,abc->kc", a, b, _structure("23").to(a.dtype))


def constraint_matrix(nt: torch.Tensor) -> torch.Tensor:
    """nt [K,9,4] (null space, columns = basis) -> [K,10,20]   (stewenius.py:82-137)."""
    e = [[nt[:, 3 * j + i] for j in range(3)] for i in range(3)]
    eet = [[2 * sum(_m11(e[i][k], e[j][k]) for k in range(3)) for j in range(3)] for i in range(3)]
    tr = eet[0][0] + eet[1][1] + eet[2][2]
    rows = []
    for i in range(3):
        for j in range(3):
            rows.append(sum(_m21(eet[i][k], e[k][j]) for k in range(3)) - 0.5 * _m21(tr, e[i][j]))
    det = (_m21(_m11(e[0][1], e[1][2]) - _m11(e[0][2], e[1][1]), e[2][0])
           + _m21(_m11(e[0][2], e[1][0]) - _m11(e[0][0], e[1][2]), e[2][1])
           + _m21(_m11(e[0][0], e[1][1]) - _m11(e[0][1], e[1][0]), e[2][2]))
    rows.append(det)
    return torch.stack(rows, dim=1)


def five_point(pts: torch.Tensor, return_aux: bool = False):
    """pts [K,5,4] -> models [K*10,3,3] (unnormalised; x2^T E x1 = 0 on the sample)."""
    K = pts.shape[0]
    A = epipolar_rows(pts)
    _, _, vh = torch.linalg.svd(A)                                    # full 9x9 Vh (:44)
    nt = vh[:, -4:, :].transpose(-1, -2)                              # [K,9,4]
    C = constraint_matrix(nt)
    elim = torch.linalg.solve(C[:, :, :10], C[:, :, 10:])
    act = torch.zeros(K, 10, 10, dtype=pts.dtype)
    act[:, 0:3] = elim[:, 0:3]
    act[:, 3] = elim[:, 4]
    act[:, 4] = elim[:, 5]
    act[:, 5] = elim[:, 7]
    act[:, 6, 0] = -1.0
    act[:, 7, 1] = -1.0
    act[:, 8, 3] = -1.0
    act[:, 9, 6] = -1.0
    ee, vv = torch.linalg.eig(act)
    E = nt.matmul(vv.real[:, -4:])                                    # [K,9,10]
    E = E.transpose(-1, -2).reshape(-1, 3, 3).transpose(-1, -2)
    if return_aux:
        return E, dict(eigvals=ee, eigvecs=vv, null=nt)
    return E
