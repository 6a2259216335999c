// Backward of the five-point solver by the implicit-function theorem.
//
// The reference differentiates through svd(A^T A), linalg.solve and eigvals with
// autograd (nister.py:117-399); svd_backward divides by differences of the four
// degenerate singular values and is only saved by the gauge invariance of the loss
// (SURVEY H2).  Here the derivative is taken at the SOLUTION instead:
//
//   G(e, P) = [ A(P) e ;  C(e) ;  (e^T e - 1)/2 ] = 0,
//
// A = 5 x 9 epipolar rows, C = the nine trace-constraint entries and det E.  With
// J = dG/de = [A; dC/de; e^T] (16 x 9, rank 9 at a regular solution)
//   de = -(J^T J)^{-1} A^T (dA e)      =>     dL/da_i = -(A v)_i e,  v = (J^T J)^{-1} dL/de,
// one 9 x 9 SPD solve per model, accumulated and factorised in `AT` (double on the
// device: J^T J squares the condition number).
#pragma once

#include "drb_common.cuh"
#include "e5_math.cuh"

namespace drb {

// In-place Cholesky solve of the symmetric positive definite n x n system S x = b
// (S full storage, row-major).  Returns false when a pivot is not positive.
template <class AT, int n>
DRB_HD bool chol_solve(AT* S, AT* b) {
    for (int j = 0; j < n; ++j) {
        AT d = S[j * n + j];
        for (int k = 0; k < j; ++k) d -= S[j * n + k] * S[j * n + k];
        if (!(d > AT(0))) return false;
        d = t_sqrt(d);
        S[j * n + j] = d;
        const AT id = AT(1) / d;
        for (int i = j + 1; i < n; ++i) {
            AT s = S[i * n + j];
            for (int k = 0; k < j; ++k) s -= S[i * n + k] * S[j * n + k];
            S[i * n + j] = s * id;
        }
    }
    for (int i = 0; i < n; ++i) {
        AT s = b[i];
        for (int k = 0; k < i; ++k) s -= S[i * n + k] * b[k];
        b[i] = s / S[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        AT s = b[i];
        for (int k = i + 1; k < n; ++k) s -= S[k * n + i] * b[k];
        b[i] = s / S[i * n + i];
    }
    return true;
}

// Jacobian of the ten cubic constraints with respect to vec(E) (row-major): Jc[10][9].
template <class AT>
DRB_HD void e5_constraint_jacobian(const AT* E, AT (*Jc)[9]) {
    AT G[9], cof[9];
    mul33_nt(E, E, G);
    const AT tr = G[0] + G[4] + G[8];
    cofactor3(E, cof);
    for (int d = 0; d < 9; ++d) {
        AT D[9];
        for (int i = 0; i < 9; ++i) D[i] = (i == d) ? AT(1) : AT(0);
        AT DEt[9], S[9], SE[9], GD[9];
        mul33_nt(D, E, DEt);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) S[3 * i + j] = DEt[3 * i + j] + DEt[3 * j + i];
        mul33(S, E, SE);
        mul33(G, D, GD);
        const AT trd = DEt[0] + DEt[4] + DEt[8];
        for (int i = 0; i < 9; ++i) Jc[i][d] = AT(2) * (SE[i] + GD[i]) - AT(2) * trd * E[i] - tr * D[i];
        Jc[9][d] = cof[d];
    }
}

// pts[j] = (x1, y1, x2, y2); E (unit norm, row-major) one solution for this sample;
// g = dL/dE.  Writes gp[j][c] = dL/d pts[j][c].
template <class T, class AT = double>
DRB_HD bool e5_backward(const T (*pts)[4], const T* E, const T* g, T (*gp)[4]) {
    AT e[9];
    for (int i = 0; i < 9; ++i) e[i] = AT(E[i]);
    AT A[5][9];
    for (int j = 0; j < 5; ++j) epipolar_row<AT>(AT(pts[j][0]), AT(pts[j][1]), AT(pts[j][2]), AT(pts[j][3]), A[j]);
    AT Jc[10][9];
    e5_constraint_jacobian<AT>(e, Jc);
    AT S[81];
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j <= i; ++j) {
            AT s = e[i] * e[j];
            for (int r = 0; r < 5; ++r) s += A[r][i] * A[r][j];
            for (int r = 0; r < 10; ++r) s += Jc[r][i] * Jc[r][j];
            S[i * 9 + j] = s;
            S[j * 9 + i] = s;
        }
    AT v[9];
    for (int i = 0; i < 9; ++i) v[i] = AT(g[i]);
    if (!chol_solve<AT, 9>(S, v)) return false;
    for (int j = 0; j < 5; ++j) {
        AT w = AT(0);
        for (int i = 0; i < 9; ++i) w += A[j][i] * v[i];
        const AT x1 = AT(pts[j][0]), y1 = AT(pts[j][1]), x2 = AT(pts[j][2]), y2 = AT(pts[j][3]);
        const AT gx1 = -w * (e[0] * x2 + e[3] * y2 + e[6]);
        const AT gy1 = -w * (e[1] * x2 + e[4] * y2 + e[7]);
        const AT gx2 = -w * (e[0] * x1 + e[1] * y1 + e[2]);
        const AT gy2 = -w * (e[3] * x1 + e[4] * y1 + e[5]);
        if (!(gx1 == gx1) || !(gy1 == gy1) || !(gx2 == gx2) || !(gy2 == gy2)) return false;
        gp[j][0] = T(gx1);
        gp[j][1] = T(gy1);
        gp[j][2] = T(gx2);
        gp[j][3] = T(gy2);
    }
    return true;
}

}  // namespace drb
