"""Oracle: Nister 5-point essential-matrix solver.

Restates `estimators/essential_matrix_estimator_nister.py:69-408` with the same
LAPACK-backed torch.linalg calls (svd of A^T A, matrix_rank filter, solve,
per-sample companion eigvals, 2x2 inverse with QR fallback), so that it is both
a numerical oracle and a fair CPU baseline.  The polynomial bookkeeping is done
with structure tensors built from exponent tuples instead of the reference's
hand-expanded `o1`/`o2`/`cs[...]` formulas; the monomial ORDER is the
reference's (`nister.py:411-422`) so intermediate stages can be compared 1:1.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import functools

import torch

# monomial orders, as exponent tuples (x, y, z)            nister.py:411-412 / 420-422
DEG1 = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
DEG2 = [(2, 0, 0), (1, 1, 0), (1, 0, 1), (1, 0, 0), (0, 2, 0), (0, 1, 1), (0, 1, 0), (0, 0, 2), (0, 0, 1),
        (0, 0, 0)]
DEG3 = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1),
        (1, 1, 0), (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2),
        (0, 0, 1), (0, 0, 0)]


@functools.lru_cache(maxsize=None)
def _structure(kind: str) -> torch.Tensor:
    """0/1 tensor T[a, b, c] = 1 iff monomial_a * monomial_b == monomial_c."""
    left, right, out = {"12": (DEG1, DEG1, DEG2), "23": (DEG2, DEG1, DEG3)}[kind]
    T = torch.zeros(len(left), len(right), len(out), dtype=torch.float64)
    for ia, ma in enumerate(left):
        for ib, mb in enumerate(right):
            mc = tuple(p + q for p, q in zip(ma, mb))
            T[ia, ib, out.index(mc)] = 1.0
    return T


def mul11(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """deg-1 x deg-1 -> deg-2 (the reference's `o1`, nister.py:410-417)."""
    return torch.einsum("ka,kb,abc->kc", a, b, _structure("12").to(a.dtype))


def mul21(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """deg-2 x deg-1 -> deg-3 (the reference's `o2`, nister.py:419-430)."""
    return torch.einsum("ka,kb,abc->kc", a, b, _structure("23").to(a.dtype))


def epipolar_rows(pts: torch.Tensor, weights=None) -> torch.Tensor:
    """A [K,s,9], row = (x1x2, x1y2, x1, y1x2, y1y2, y1, x2, y2, 1)  (nister.py:84-115)."""
    x1, y1, x2, y2 = pts[..., 0], pts[..., 1], pts[..., 2], pts[..., 3]
    A = torch.stack((x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, torch.ones_like(x1)), dim=-1)
    if weights is not None:
        A = weights.unsqueeze(-1) * A
    return A


def null_space(A: torch.Tensor) -> torch.Tensor:
    """Last four rows of Vh of A^T A (nister.py:117-119) -> [K,4,9]."""
    _, _, vh = torch.linalg.svd(A.transpose(-1, -2) @ A)
    return vh[:, -4:, :]


def constraint_matrix(null: torch.Tensor) -> torch.Tensor:
    """[K,10,20]: rows 0..8 = entries (i,j) of E E^T E - 0.5 tr(E E^T) E, row 9 =
    det E, as cubic polynomials in (x,y,z) with E = x n0 + y n1 + z n2 + n3
    (nister.py:121-152)."""
    K = null.shape[0]
    nt = null.transpose(-1, -2)                       # [K,9,4]; entry (i,j) of E is vec index 3j+i
    e = [[nt[:, 3 * j + i] for j in range(3)] for i in range(3)]
    # E E^T (symmetric), then subtract half the trace on the diagonal  (nister.py:132-143)
    eet = [[sum(mul11(e[i][k], e[j][k]) for k in range(3)) for j in range(3)] for i in range(3)]
    half_tr = 0.5 * (eet[0][0] + eet[1][1] + eet[2][2])
    lam = [[eet[i][j] - (half_tr if i == j else 0) for j in range(3)] for i in range(3)]
    rows = []
    for i in range(3):
        for j in range(3):
            rows.append(sum(mul21(lam[i][k], e[k][j]) for k in range(3)))       # nister.py:145-152
    det = (mul21(mul11(e[0][1], e[1][2]) - mul11(e[0][2], e[1][1]), e[2][0])
           + mul21(mul11(e[0][2], e[1][0]) - mul11(e[0][0], e[1][2]), e[2][1])
           + mul21(mul11(e[0][0], e[1][1]) - mul11(e[0][1], e[1][0]), e[2][2]))  # nister.py:126-128
    rows.append(det)
    return torch.stack(rows, dim=1).reshape(K, 10, 20)


def rank_filter(coeffs: torch.Tensor) -> torch.Tensor:
    """nister.py:155-157."""
    r_left = torch.linalg.matrix_rank(coeffs[:, :, :10])
    r_all = torch.linalg.matrix_rank(coeffs)
    return r_left >= torch.max(r_all, torch.ones_like(r_left) * 10)


def z_polynomial_matrix(elim: torch.Tensor) -> torch.Tensor:
    """[K',3,13] from the eliminated right block [K',10,10] (nister.py:165-176).
    Row i = (row 4+2i) - z * (row 5+2i); layout per row: cubic(z) multiplying x
    (cols 0-3), cubic(z) multiplying y (cols 4-7), quartic(z) (cols 8-12), all
    with DESCENDING powers of z."""
    Kp = elim.shape[0]
    A = torch.zeros(Kp, 3, 13, dtype=elim.dtype)
    for i in range(3):
        ra, rb = elim[:, 4 + 2 * i], elim[:, 5 + 2 * i]
        for off, lo, n in ((0, 0, 3), (4, 3, 3), (8, 6, 4)):
            A[:, i, off + 1: off + 1 + n] += ra[:, lo: lo + n]
            A[:, i, off: off + n] -= rb[:, lo: lo + n]
    return A


def _polymul_asc(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(a.shape[0], a.shape[1] + b.shape[1] - 1, dtype=a.dtype)
    for i in range(a.shape[1]):
        out[:, i: i + b.shape[1]] += a[:, i: i + 1] * b
    return out


def determinant_polynomial(A: torch.Tensor) -> torch.Tensor:
    """cs [K',11], cs[i] = coefficient of z^i of det [cx | cy | cq]  (nister.py:178-348)."""
    cx = A[:, :, 0:4].flip(-1)      # ascending powers
    cy = A[:, :, 4:8].flip(-1)
    cq = A[:, :, 8:13].flip(-1)

    def minor(r, s):
        return _polymul_asc(cx[:, r], cy[:, s]) - _polymul_asc(cx[:, s], cy[:, r])

    return (_polymul_asc(minor(1, 2), cq[:, 0]) - _polymul_asc(minor(0, 2), cq[:, 1])
            + _polymul_asc(minor(0, 1), cq[:, 2]))


def _models_from_roots(A_i, null_i, roots):
    """nister.py:379-401 for one sample: back-substitute (x, y) for each root z."""
    r1 = roots
    r2, r3, r4 = r1 * r1, r1 ** 3, r1 ** 4
    bx = A_i[:, 0:1] * r3 + A_i[:, 1:2] * r2 + A_i[:, 2:3] * r1 + A_i[:, 3:4]          # [3,R]
    by = A_i[:, 4:5] * r3 + A_i[:, 5:6] * r2 + A_i[:, 6:7] * r1 + A_i[:, 7:8]
    Bs = torch.stack((bx, by), dim=0).transpose(0, -1)                                # [R,3,2]
    bs = (A_i[:, 8:9] * r4 + A_i[:, 9:10] * r3 + A_i[:, 10:11] * r2 + A_i[:, 11:12] * r1
          + A_i[:, 12:13]).T.unsqueeze(-1)                                             # [R,3,1]
    xy = torch.linalg.inv(Bs[:, 0:2, 0:2]) @ bs[:, 0:2]
    bad = ((Bs[:, 2].unsqueeze(1) @ xy - bs[:, 2].unsqueeze(1)).abs() > 1e-3).flatten()
    if bad.any():
        q, r = torch.linalg.qr(Bs[bad])
        xy[bad] = torch.linalg.solve(r, q.transpose(-1, -2) @ bs[bad])
    Es = null_i[0] * (-xy[:, 0]) + null_i[1] * (-xy[:, 1]) + null_i[2] * roots.unsqueeze(-1) + null_i[3]
    inv = 1.0 / torch.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2 + roots.unsqueeze(-1) ** 2 + 1.0)
    return Es * inv


def five_point(pts: torch.Tensor, weights=None, return_aux: bool = False):
    """pts [K,5,4] -> models [K'*10,3,3] with x2^T E x1 = 0, ||E||_F = 1.

    Every sample that survives the rank filter emits exactly ten slots: one per
    companion-matrix eigenvalue, REAL PART taken for complex ones (nister.py:370,
    SURVEY D3).  With `return_aux` also returns a dict(keep, null, coeffs, A, cs,
    roots[K',10] complex) used by the stage-level tests.
    """
    A_s = epipolar_rows(pts, weights)
    null = null_space(A_s)
    coeffs = constraint_matrix(null)
    keep = rank_filter(coeffs)
    coeffs_k = coeffs[keep]
    elim = torch.linalg.solve(coeffs_k[:, :, :10], coeffs_k[:, :, 10:])
    A = z_polynomial_matrix(elim)
    cs = determinant_polynomial(A)
    null_k = null[keep]
    models, all_roots = [], []
    for bi in range(A.shape[0]):                                   # nister.py:355 -- the K-trip python loop
        C = torch.zeros(10, 10, dtype=cs.dtype)
        C[:-1, 1:] = torch.eye(9, dtype=cs.dtype)
        C[-1] = -cs[bi, :-1] / cs[bi, -1]
        if not torch.isfinite(C).all():
            continue
        ev = torch.linalg.eigvals(C)
        all_roots.append(ev)
        models.append(_models_from_roots(A[bi], null_k[bi], ev.real))
    if not models:
        out = torch.eye(3, dtype=cs.dtype).unsqueeze(0)
    else:
        out = torch.cat(models).view(-1, 3, 3).transpose(-1, -2)
    if return_aux:
        aux = dict(keep=keep, null=null, coeffs=coeffs, A=A, cs=cs,
                   roots=torch.stack(all_roots) if all_roots else torch.zeros(0, 10, dtype=torch.complex64))
        return out, aux
    return out


def real_root_mask(roots: torch.Tensor, tol: float = 1e-4) -> torch.Tensor:
    """Slots whose companion eigenvalue is real: the only slots on which the
    reference's output is a genuine essential matrix."""
    return roots.imag.abs() <= tol * (1.0 + roots.real.abs())
