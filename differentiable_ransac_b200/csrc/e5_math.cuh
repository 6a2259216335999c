// Five-point essential-matrix solver (Nister's formulation), one hypothesis per
// thread.  Replaces estimators/essential_matrix_estimator_nister.py:69-408
// (and serves estimators/essential_matrix_estimator_stewenius.py:20-80, which
// solves the same polynomial system through an action matrix).
//
// Differences from the reference's arithmetic, all inside the reference's own
// noise floor (DESIGN.md "Parity contract"):
//   * the 4-dim null space comes from a Householder QR of A^T instead of an SVD
//     of A^T A -- any orthonormal basis spans the same space, and the reference's
//     basis is itself LAPACK-rounding dependent;
//   * only REAL roots of the degree-10 polynomial yield models; the reference
//     keeps Re(z) of complex eigenvalues, which are not essential matrices
//     (SURVEY D3).  Unused slots hold the identity with valid = 0, like the
//     reference's padding (nister.py:400-401);
//   * each root is polished by Gauss-Newton on the ten cubic constraints
//     evaluated from E directly, which removes the elimination's rounding error.
//
// vec(E) is ROW-major: E[i][j] = e[3 i + j], constraint x2^T E x1 = 0.
#pragma once

#include "drb_common.cuh"
#include "poly_gen.cuh"
#include "poly_roots.cuh"

namespace drb {

// Epipolar row for vec(E) row-major: (x2 x1, x2 y1, x2, y2 x1, y2 y1, y2, x1, y1, 1)
template <class T>
DRB_HD void epipolar_row(T x1, T y1, T x2, T y2, T* r) {
    r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2;
    r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2;
    r[6] = x1;      r[7] = y1;      r[8] = T(1);
}

// Householder QR of A^T (9 x R, column k = row k of A), in place: the reflector vectors stay in the lower
// trapezoid of W, their scales in beta.  R in {5, 7, 8}.
template <class T, int R>
DRB_HD void householder_factor(const T (*rows)[9], T (*W)[R], T* beta) {
    DRB_UNROLL
    for (int k = 0; k < R; ++k) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) W[i][k] = rows[k][i];
    }
    DRB_UNROLL
    for (int k = 0; k < R; ++k) {
        T nrm2 = T(0);
        DRB_UNROLL
        for (int i = k; i < 9; ++i) nrm2 += W[i][k] * W[i][k];
        const T nrm = t_sqrt(nrm2);
        const T alpha = W[k][k] > T(0) ? -nrm : nrm;
        W[k][k] -= alpha;  // v = x - alpha e_k, stored in place
        T vtv = T(0);
        DRB_UNROLL
        for (int i = k; i < 9; ++i) vtv += W[i][k] * W[i][k];
        beta[k] = vtv > T(0) ? T(2) / vtv : T(0);
        DRB_UNROLL
        for (int j = k + 1; j < R; ++j) {
            T dot = T(0);
            DRB_UNROLL
            for (int i = k; i < 9; ++i) dot += W[i][k] * W[i][j];
            dot *= beta[k];
            DRB_UNROLL
            for (int i = k; i < 9; ++i) W[i][j] -= dot * W[i][k];
        }
    }
}

// Null vector m (0 <= m < 9 - R) of the factored matrix: Q e_{R+m} = H_0 H_1 ... H_{R-1} e_{R+m}.  `m` may be a
// run-time value (the lanes of a cooperative group take one vector each): it only selects the unit vector.
template <class T, int R>
DRB_HD void householder_null_vector(const T (*W)[R], const T* beta, int m, T* q) {
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) q[i] = (i == R + m) ? T(1) : T(0);
    DRB_UNROLL
    for (int k = R - 1; k >= 0; --k) {
        T dot = T(0);
        DRB_UNROLL
        for (int i = k; i < 9; ++i) dot += W[i][k] * q[i];
        dot *= beta[k];
        DRB_UNROLL
        for (int i = k; i < 9; ++i) q[i] -= dot * W[i][k];
    }
}

// Orthonormal basis of the null space of the R x 9 matrix whose rows are given
// (R in {5, 7, 8}) by Householder QR of its transpose.  Writes 9 - R vectors.
template <class T, int R>
DRB_HD void null_space_rows(const T (*rows)[9], T (*null)[9]) {
    T W[9][R];
    T beta[R];
    householder_factor<T, R>(rows, W, beta);
    DRB_UNROLL
    for (int mI = 0; mI < 9 - R; ++mI) householder_null_vector<T, R>(W, beta, mI, null[mI]);
}

// Fill the 10 x 20 constraint matrix (rows 0..8: E E^T E - 1/2 tr(E E^T) E,
// row 9: det E) for E = x N0 + y N1 + z N2 + N3.  `M(r, c)` is the per-thread
// scratch accessor (shared memory on the device).
template <class T, class Mat>
DRB_HD void e5_constraints(const T (*N)[9], Mat& M) {
    T e[9][4];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) {
        DRB_UNROLL
        for (int b = 0; b < 4; ++b) e[i][b] = N[b][i];
    }
    // Lambda = E E^T - 1/2 tr(E E^T) I, symmetric: indices (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
    T L[6][10];
    {
        int s = 0;
        DRB_UNROLL
        for (int i = 0; i < 3; ++i) {
            DRB_UNROLL
            for (int j = i; j < 3; ++j) {
                poly_mul11(e[3 * i + 0], e[3 * j + 0], L[s]);
                poly_mul11_acc(e[3 * i + 1], e[3 * j + 1], L[s]);
                poly_mul11_acc(e[3 * i + 2], e[3 * j + 2], L[s]);
                ++s;
            }
        }
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) {
            const T ht = T(0.5) * (L[0][c] + L[3][c] + L[5][c]);
            L[0][c] -= ht;
            L[3][c] -= ht;
            L[5][c] -= ht;
        }
    }
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) {
            T row[20];
            DRB_UNROLL
            for (int c = 0; c < 20; ++c) row[c] = T(0);
            DRB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const int lo = i < k ? i : k, hi = i < k ? k : i;
                const int s = (lo == 0) ? hi : (lo == 1 ? 2 + hi : 5);
                poly_mul21_acc(L[s], e[3 * k + j], row);
            }
            DRB_UNROLL
            for (int c = 0; c < 20; ++c) M(3 * i + j, c) = row[c];
        }
    }
    {   // determinant, expanded along the third row
        T q[10], q2[10], row[20];
        DRB_UNROLL
        for (int c = 0; c < 20; ++c) row[c] = T(0);
        poly_mul11(e[1], e[5], q);
        poly_mul11(e[2], e[4], q2);
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) q[c] -= q2[c];
        poly_mul21_acc(q, e[6], row);
        poly_mul11(e[2], e[3], q);
        poly_mul11(e[0], e[5], q2);
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) q[c] -= q2[c];
        poly_mul21_acc(q, e[7], row);
        poly_mul11(e[0], e[4], q);
        poly_mul11(e[1], e[3], q2);
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) q[c] -= q2[c];
        poly_mul21_acc(q, e[8], row);
        DRB_UNROLL
        for (int c = 0; c < 20; ++c) M(9, c) = row[c];
    }
}

// Gauss-Jordan on the left 10 x 10 block with partial pivoting; rows 4..9 of the
// right block are fully reduced (that is all the z-polynomials need).  Returns
// false when a pivot is negligible (the reference drops such samples through its
// matrix_rank filter, nister.py:155-157).
template <class T, class Mat>
DRB_HD bool e5_eliminate(Mat& M) {
    T scale = T(0);
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 20; ++c) scale = t_max(scale, t_abs(M(r, c)));
    const T tol = scale * (sizeof(T) == 4 ? T(1e-6) : T(1e-13));
    bool ok = true;
    for (int p = 0; p < 10; ++p) {
        int piv = p;
        T best = t_abs(M(p, p));
        for (int r = p + 1; r < 10; ++r) {
            const T v = t_abs(M(r, p));
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > tol)) ok = false;
        T prow[20];
        const T ip = t_rcp(M(piv, p));
        DRB_UNROLL
        for (int c = 0; c < 20; ++c) {
            const T a = M(piv, c);
            const T b = M(p, c);
            M(piv, c) = b;          // swap rows p <-> piv
            prow[c] = a * ip;       // normalised pivot row
            M(p, c) = prow[c];
        }
        for (int r = p + 1; r < 10; ++r) {
            const T f = M(r, p);
            DRB_UNROLL
            for (int c = 0; c < 20; ++c)
                if (c > p) M(r, c) -= f * prow[c];
        }
    }
    // back-substitution restricted to rows 4..8, right block only
    for (int p = 9; p >= 5; --p) {
        T prow[10];
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) prow[c] = M(p, 10 + c);
        for (int r = 4; r < p; ++r) {
            const T f = M(r, p);
            DRB_UNROLL
            for (int c = 0; c < 10; ++c) M(r, 10 + c) -= f * prow[c];
        }
    }
    return ok;
}

// Gauss-Newton polish of (x, y, z) on the ten cubic constraints evaluated from
// E = x N0 + y N1 + z N2 + N3 directly.
template <class T>
DRB_HD void e5_polish(const T (*N)[9], T& x, T& y, T& z, int iters) {
    for (int it = 0; it < iters; ++it) {
        T E[9];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) E[i] = x * N[0][i] + y * N[1][i] + z * N[2][i] + N[3][i];
        T G[9], GE[9], cof[9];
        mul33_nt(E, E, G);
        const T tr = G[0] + G[4] + G[8];
        mul33(G, E, GE);
        cofactor3(E, cof);
        T r[10];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) r[i] = T(2) * GE[i] - tr * E[i];
        r[9] = det3(E);
        T JtJ[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        T Jtr[3] = {T(0), T(0), T(0)};
        T J[3][10];
        DRB_UNROLL
        for (int d = 0; d < 3; ++d) {
            const T* D = N[d];
            T DEt[9], S[9], SE[9], GD[9];
            mul33_nt(D, E, DEt);
            DRB_UNROLL
            for (int i = 0; i < 3; ++i) {
                DRB_UNROLL
                for (int j = 0; j < 3; ++j) S[3 * i + j] = DEt[3 * i + j] + DEt[3 * j + i];
            }
            mul33(S, E, SE);
            mul33(G, D, GD);
            const T trd = DEt[0] + DEt[4] + DEt[8];
            T dd = T(0);
            DRB_UNROLL
            for (int i = 0; i < 9; ++i) {
                J[d][i] = T(2) * (SE[i] + GD[i]) - T(2) * trd * E[i] - tr * D[i];
                dd += cof[i] * D[i];
            }
            J[d][9] = dd;
        }
        DRB_UNROLL
        for (int i = 0; i < 10; ++i) {
            JtJ[0] += J[0][i] * J[0][i];
            JtJ[1] += J[0][i] * J[1][i];
            JtJ[2] += J[0][i] * J[2][i];
            JtJ[3] += J[1][i] * J[1][i];
            JtJ[4] += J[1][i] * J[2][i];
            JtJ[5] += J[2][i] * J[2][i];
            Jtr[0] += J[0][i] * r[i];
            Jtr[1] += J[1][i] * r[i];
            Jtr[2] += J[2][i] * r[i];
        }
        // solve the symmetric 3x3 system by cofactors
        const T a = JtJ[0], b = JtJ[1], c = JtJ[2], d = JtJ[3], e = JtJ[4], f = JtJ[5];
        const T c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
        const T det = a * c00 + b * c01 + c * c02;
        if (!(t_abs(det) > T(0))) return;
        const T c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
        const T id = t_rcp(det);
        const T dx = -(c00 * Jtr[0] + c01 * Jtr[1] + c02 * Jtr[2]) * id;
        const T dy = -(c01 * Jtr[0] + c11 * Jtr[1] + c12 * Jtr[2]) * id;
        const T dz = -(c02 * Jtr[0] + c12 * Jtr[1] + c22 * Jtr[2]) * id;
        if (!(dx == dx) || !(dy == dy) || !(dz == dz)) return;
        // trust region: never move further than the current parameter scale
        const T step2 = dx * dx + dy * dy + dz * dz;
        const T lim2 = T(0.25) * (x * x + y * y + z * z + T(1));
        if (step2 > lim2) return;
        x += dx;
        y += dy;
        z += dz;
    }
}

// Plain-array sink for e5_solve's models (host harness, and anything that wants [10][9]).
template <class T>
struct ArrayModelSink {
    T (*m)[9];
    DRB_HD void operator()(int slot, int i, T v) { m[slot][i] = v; }
};

// The per-sample data every root of the sample needs: the null-space basis, the three z-polynomial
// equations cx_i x + cy_i y + cq_i = 0 (ascending powers) and their determinant polynomial P.
// 86 scalars: the device kernel parks them in the thread's scratch column so that ANY lane of the warp
// can turn one of this sample's root brackets into a model (solver_e5.cu).
template <class T>
struct E5Sample {
    T N[4][9];
    T cx[3][4], cy[3][4], cq[3][5];
    T P[11];
};
static constexpr int kE5SampleScalars = 36 + 12 + 12 + 15 + 11;

// Stage 1b: S.N (a basis of the 4-dimensional space E lives in) -> the rest of E5Sample.  `M` is the
// 10 x 20 scratch; it is dead when this returns.  Shared by the minimal solver (null space of the 5 x 9
// system) and the non-minimal refit (the 4 smallest eigenvectors of A^T A, refit_math.cuh).
template <class T, class Mat>
DRB_HD bool e5_prepare_from_null(Mat& M, E5Sample<T>& S) {
    e5_constraints<T, Mat>(S.N, M);
    bool ok = e5_eliminate<T, Mat>(M);
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        T ra[10], rb[10];
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) {
            ra[c] = M(4 + 2 * i, 10 + c);
            rb[c] = M(5 + 2 * i, 10 + c);
        }
        S.cx[i][0] = ra[2]; S.cx[i][1] = ra[1] - rb[2]; S.cx[i][2] = ra[0] - rb[1]; S.cx[i][3] = -rb[0];
        S.cy[i][0] = ra[5]; S.cy[i][1] = ra[4] - rb[5]; S.cy[i][2] = ra[3] - rb[4]; S.cy[i][3] = -rb[3];
        S.cq[i][0] = ra[9]; S.cq[i][1] = ra[8] - rb[9]; S.cq[i][2] = ra[7] - rb[8]; S.cq[i][3] = ra[6] - rb[7];
        S.cq[i][4] = -rb[6];
    }
    // determinant polynomial (degree 10) of [cx | cy | cq]
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) S.P[i] = T(0);
    DRB_UNROLL
    for (int t = 0; t < 3; ++t) {
        const int r = (t == 0) ? 1 : 0;
        const int s = (t == 2) ? 1 : 2;
        const T sign = (t == 1) ? T(-1) : T(1);
        T mn[7];
        DRB_UNROLL
        for (int i = 0; i < 7; ++i) mn[i] = T(0);
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) {
            DRB_UNROLL
            for (int j = 0; j < 4; ++j) mn[i + j] += S.cx[r][i] * S.cy[s][j] - S.cx[s][i] * S.cy[r][j];
        }
        DRB_UNROLL
        for (int i = 0; i < 7; ++i) {
            DRB_UNROLL
            for (int j = 0; j < 5; ++j) S.P[i + j] += sign * mn[i] * S.cq[t][j];
        }
    }
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) ok = ok && (S.P[i] == S.P[i]) && (t_abs(S.P[i]) < T(1e30));
    return ok;
}

// Stage 1: points -> E5Sample.
template <class T, class Mat>
DRB_HD bool e5_prepare(const T (*pts)[4], Mat& M, E5Sample<T>& S) {
    {
        T rows[5][9];
        DRB_UNROLL
        for (int j = 0; j < 5; ++j) epipolar_row(pts[j][0], pts[j][1], pts[j][2], pts[j][3], rows[j]);
        null_space_rows<T, 5>(rows, S.N);
    }
    return e5_prepare_from_null<T, Mat>(M, S);
}

// Stage 3: one root z of P -> (x, y) from the best-conditioned pair of the three equations -> Gauss-Newton
// polish -> unit-norm E.  Returns false when the root has to be dropped.
template <class T>
DRB_HD bool e5_model_from_root(const E5Sample<T>& S, T z, int polish_iters, T* E) {
    T vx[3], vy[3], vq[3];
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        vx[i] = ((S.cx[i][3] * z + S.cx[i][2]) * z + S.cx[i][1]) * z + S.cx[i][0];
        vy[i] = ((S.cy[i][3] * z + S.cy[i][2]) * z + S.cy[i][1]) * z + S.cy[i][0];
        vq[i] = (((S.cq[i][4] * z + S.cq[i][3]) * z + S.cq[i][2]) * z + S.cq[i][1]) * z + S.cq[i][0];
    }
    const T d01 = vx[0] * vy[1] - vx[1] * vy[0];
    const T d02 = vx[0] * vy[2] - vx[2] * vy[0];
    const T d12 = vx[1] * vy[2] - vx[2] * vy[1];
    T x, y;
    if (t_abs(d01) >= t_abs(d02) && t_abs(d01) >= t_abs(d12)) {
        const T id = t_rcp(d01);
        x = (vq[1] * vy[0] - vq[0] * vy[1]) * id;
        y = (vq[0] * vx[1] - vq[1] * vx[0]) * id;
    } else if (t_abs(d02) >= t_abs(d12)) {
        const T id = t_rcp(d02);
        x = (vq[2] * vy[0] - vq[0] * vy[2]) * id;
        y = (vq[0] * vx[2] - vq[2] * vx[0]) * id;
    } else {
        const T id = t_rcp(d12);
        x = (vq[2] * vy[1] - vq[1] * vy[2]) * id;
        y = (vq[1] * vx[2] - vq[2] * vx[1]) * id;
    }
    if (!(x == x) || !(y == y) || t_abs(x) > T(1e18) || t_abs(y) > T(1e18)) return false;
    e5_polish<T>(S.N, x, y, z, polish_iters);
    T n2 = T(0);
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) {
        E[i] = x * S.N[0][i] + y * S.N[1][i] + z * S.N[2][i] + S.N[3][i];
        n2 += E[i] * E[i];
    }
    if (!(n2 > T(0)) || !(n2 < T(1e37))) return false;
    const T inv = t_rsqrt(n2);
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) E[i] *= inv;
    return true;
}

// Full solver, serial composition of the three stages (host harness; the device kernel runs stage 3
// warp-cooperatively instead).  `models(slot, i, v)` receives the row-major unit-norm solutions in slots
// 0..n-1 (n = return value; slots >= n are NOT written: the caller pads with the identity).
template <class T, class Mat, class RT, class Sink>
DRB_HD int e5_solve(const T (*pts)[4], Mat& M, Sink& models, int polish_iters = 2) {
    E5Sample<T> S;
    if (!e5_prepare<T, Mat>(pts, M, S)) return 0;
    RT Pr[11], blo[10], bhi[10];
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) Pr[i] = RT(S.P[i]);
    int n0 = 0;
    const int nb = isolate_deg10<RT>(Pr, blo, bhi, n0);
    int nout = 0;
    for (int r = 0; r < nb; ++r) {
        RT zr;
        if (!root_from_bracket<RT>(Pr, r >= n0, blo[r], bhi[r], zr)) continue;
        T E[9];
        if (!e5_model_from_root<T>(S, T(zr), polish_iters, E)) continue;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) models(nout, i, E[i]);
        ++nout;
    }
    return nout;
}

}  // namespace drb
