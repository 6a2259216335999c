// Normalised 8-point fundamental matrix (forward + implicit-function backward) and a
// correct 7-point solver, one hypothesis per thread.
//
// Replaces FundamentalMatrixEstimatorNew.normalize / estimate_non_minimal_model
// (estimators/fundamental_matrix_estimator.py:177-260).  The reference takes the last
// right singular vector of A^T A; the null vector of the 8 x 9 matrix A is the same
// line, obtained here by a Householder QR of A^T (unit norm, sign arbitrary -- as is
// LAPACK's).  No rank-2 projection and no rescaling after de-normalisation, like the
// reference (:254-259).  The 7-point path of the reference is broken, so
// f7_solve follows the textbook algorithm and is validated by its own algebra.
#pragma once

#include "drb_common.cuh"
#include "e5_math.cuh"      // epipolar_row, null_space_rows
#include "e5_backward.cuh"  // (shares nothing but keeps include order stable)

namespace drb {

template <class T>
struct HartleyNorm {
    T m[4];    // mass point (x1, y1, x2, y2)
    T r1, r2;  // sqrt(2) / mean distance to the mass point
};

template <class T, int S>
DRB_HD HartleyNorm<T> hartley_normalize(const T (*pts)[4], T (*npts)[4]) {
    HartleyNorm<T> h;
    DRB_UNROLL
    for (int c = 0; c < 4; ++c) {
        T s = T(0);
        DRB_UNROLL
        for (int j = 0; j < S; ++j) s += pts[j][c];
        h.m[c] = s / T(S);
    }
    T d1 = T(0), d2 = T(0);
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        DRB_UNROLL
        for (int c = 0; c < 4; ++c) npts[j][c] = pts[j][c] - h.m[c];
        d1 += t_sqrt(npts[j][0] * npts[j][0] + npts[j][1] * npts[j][1]);
        d2 += t_sqrt(npts[j][2] * npts[j][2] + npts[j][3] * npts[j][3]);
    }
    const T sqrt2 = T(1.4142135623730951);
    h.r1 = sqrt2 / (d1 / T(S));
    h.r2 = sqrt2 / (d2 / T(S));
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        npts[j][0] *= h.r1; npts[j][1] *= h.r1;
        npts[j][2] *= h.r2; npts[j][3] *= h.r2;
    }
    return h;
}

// F = T2^T Fn T1 with T = [[r, 0, -r mx], [0, r, -r my], [0, 0, 1]]
template <class T>
DRB_HD void denormalize_f(const T* Fn, const HartleyNorm<T>& h, T* F) {
    T G[9];
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        G[3 * i + 0] = Fn[3 * i + 0] * h.r1;
        G[3 * i + 1] = Fn[3 * i + 1] * h.r1;
        G[3 * i + 2] = Fn[3 * i + 2] - h.r1 * (h.m[0] * Fn[3 * i + 0] + h.m[1] * Fn[3 * i + 1]);
    }
    DRB_UNROLL
    for (int j = 0; j < 3; ++j) {
        F[j] = h.r2 * G[j];
        F[3 + j] = h.r2 * G[3 + j];
        F[6 + j] = G[6 + j] - h.r2 * (h.m[2] * G[j] + h.m[3] * G[3 + j]);
    }
}

// pts[8][4] -> F[9]; returns false if the result is not finite.
template <class T>
DRB_HD bool f8_solve(const T (*pts)[4], T* F) {
    T n[8][4];
    const HartleyNorm<T> h = hartley_normalize<T, 8>(pts, n);
    T rows[8][9];
    DRB_UNROLL
    for (int j = 0; j < 8; ++j) epipolar_row(n[j][0], n[j][1], n[j][2], n[j][3], rows[j]);
    T f[1][9];
    null_space_rows<T, 8>(rows, f);
    denormalize_f(f[0], h, F);
    bool ok = true;
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) ok = ok && (F[i] == F[i]) && (t_abs(F[i]) < T(1e30));
    return ok;
}

// Solve the general n x n system J x = b in place by Gauss-Jordan with partial pivoting.
template <class AT, int n>
DRB_HD bool gauss_solve(AT* J, AT* b) {
    for (int p = 0; p < n; ++p) {
        int piv = p;
        AT best = t_abs(J[p * n + p]);
        for (int r = p + 1; r < n; ++r) {
            const AT v = t_abs(J[r * n + p]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > AT(0))) return false;
        if (piv != p) {
            for (int c = 0; c < n; ++c) { const AT t = J[p * n + c]; J[p * n + c] = J[piv * n + c]; J[piv * n + c] = t; }
            const AT t = b[p]; b[p] = b[piv]; b[piv] = t;
        }
        const AT ip = AT(1) / J[p * n + p];
        for (int r = 0; r < n; ++r) {
            if (r == p) continue;
            const AT f = J[r * n + p] * ip;
            if (f == AT(0)) continue;
            for (int c = p; c < n; ++c) J[r * n + c] -= f * J[p * n + c];
            b[r] -= f * b[p];
        }
    }
    for (int p = 0; p < n; ++p) b[p] /= J[p * n + p];
    return true;
}

// Solve the symmetric positive definite 8 x 8 system M u = b in place (Cholesky, no pivoting; every loop has
// constant bounds so the whole factorisation stays in registers).  Returns false on a non-positive pivot.
template <class AT>
DRB_HD bool chol_solve8(AT (*M)[8], AT* b) {
    DRB_UNROLL
    for (int j = 0; j < 8; ++j) {
        AT d = M[j][j];
        DRB_UNROLL
        for (int k = 0; k < 8; ++k)
            if (k < j) d -= M[j][k] * M[j][k];
        if (!(d > AT(0))) return false;
        const AT inv = AT(1) / t_sqrt(d);
        M[j][j] = d * inv;  // = sqrt(d)
        DRB_UNROLL
        for (int i = 0; i < 8; ++i)
            if (i > j) {
                AT s = M[i][j];
                DRB_UNROLL
                for (int k = 0; k < 8; ++k)
                    if (k < j) s -= M[i][k] * M[j][k];
                M[i][j] = s * inv;
            }
    }
    DRB_UNROLL
    for (int i = 0; i < 8; ++i) {  // L y = b
        AT s = b[i];
        DRB_UNROLL
        for (int k = 0; k < 8; ++k)
            if (k < i) s -= M[i][k] * b[k];
        b[i] = s / M[i][i];
    }
    DRB_UNROLL
    for (int ii = 0; ii < 8; ++ii) {  // L^T u = y
        const int i = 7 - ii;
        AT s = b[i];
        DRB_UNROLL
        for (int k = 0; k < 8; ++k)
            if (k > i) s -= M[k][i] * b[k];
        b[i] = s / M[i][i];
    }
    return true;
}

// Backward of f8_solve: g = dL/dF -> gp[8][4] = dL/dpts.
// Fn = unit null vector of A(n):  [A; f^T] df = [-dA f; 0]  =>  dL/dA = -u f^T with
// [A; f^T]^T [u; lambda] = dL/dFn.  Because A f = 0 the bordered system splits: lambda = f . dL/dFn and
// A^T u = dL/dFn - lambda f, solved through the 8 x 8 normal equations (A A^T) u = A (dL/dFn - lambda f) in AT
// (double: cond(A)^2 ~ 1e8 leaves eight digits).  Then the chain through the Hartley normalisation.
// `F_fwd` (nullable) is the forward's output: it fixes the sign of the recomputed null vector; without it the
// forward is re-run in T for that purpose.
template <class T, class AT = double>
DRB_HD bool f8_backward(const T (*pts)[4], const T* g, T (*gp)[4], const T* F_fwd = nullptr) {
    AT p[8][4], n[8][4];
    for (int j = 0; j < 8; ++j)
        for (int c = 0; c < 4; ++c) p[j][c] = AT(pts[j][c]);
    const HartleyNorm<AT> h = hartley_normalize<AT, 8>(p, n);
    AT rows[8][9];
    for (int j = 0; j < 8; ++j) epipolar_row(n[j][0], n[j][1], n[j][2], n[j][3], rows[j]);
    AT fv[1][9];
    null_space_rows<AT, 8>(rows, fv);
    const AT* f = fv[0];
    if (F_fwd != nullptr) {
        AT Fd[9], dot = AT(0);
        denormalize_f(f, h, Fd);
        for (int i = 0; i < 9; ++i) dot += AT(F_fwd[i]) * Fd[i];
        if (dot < AT(0))
            for (int i = 0; i < 9; ++i) fv[0][i] = -fv[0][i];
    } else {
        // the forward ran in T: align the sign of the recomputed null vector with it
        T nf[8][4];
        const HartleyNorm<T> hf = hartley_normalize<T, 8>(pts, nf);
        (void)hf;
        T rf[8][9];
        for (int j = 0; j < 8; ++j) epipolar_row(nf[j][0], nf[j][1], nf[j][2], nf[j][3], rf[j]);
        T ff[1][9];
        null_space_rows<T, 8>(rf, ff);
        AT dot = AT(0);
        for (int i = 0; i < 9; ++i) dot += AT(ff[0][i]) * f[i];
        if (dot < AT(0))
            for (int i = 0; i < 9; ++i) fv[0][i] = -fv[0][i];
    }
    AT gF[9];
    for (int i = 0; i < 9; ++i) gF[i] = AT(g[i]);
    // T1, T2 and intermediate products
    const AT T1[9] = {h.r1, 0, -h.r1 * h.m[0], 0, h.r1, -h.r1 * h.m[1], 0, 0, 1};
    const AT T2[9] = {h.r2, 0, -h.r2 * h.m[2], 0, h.r2, -h.r2 * h.m[3], 0, 0, 1};
    AT G[9], H[9], tmp[9], gFn[9], gT1[9], gT2[9];
    mul33(f, T1, G);        // G = Fn T1
    mul33_tn(T2, f, H);     // H = T2^T Fn
    mul33(T2, gF, tmp);     // dL/dFn = T2 g T1^T
    mul33_nt(tmp, T1, gFn);
    mul33_tn(H, gF, gT1);   // dL/dT1 = H^T g
    mul33_nt(G, gF, gT2);   // dL/dT2 = G g^T
    AT g_r1 = gT1[0] + gT1[4] - h.m[0] * gT1[2] - h.m[1] * gT1[5];
    AT g_r2 = gT2[0] + gT2[4] - h.m[2] * gT2[2] - h.m[3] * gT2[5];
    AT g_m[4] = {-h.r1 * gT1[2], -h.r1 * gT1[5], -h.r2 * gT2[2], -h.r2 * gT2[5]};
    // adjoint of the null vector
    AT v[8];
    {
        AT lam = AT(0), rhs[9], MM[8][8];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) lam += f[i] * gFn[i];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) rhs[i] = gFn[i] - lam * f[i];
        DRB_UNROLL
        for (int r = 0; r < 8; ++r) {
            AT s = AT(0);
            DRB_UNROLL
            for (int i = 0; i < 9; ++i) s += rows[r][i] * rhs[i];
            v[r] = s;
            DRB_UNROLL
            for (int c = 0; c < 8; ++c)
                if (c <= r) {
                    AT m = AT(0);
                    DRB_UNROLL
                    for (int i = 0; i < 9; ++i) m += rows[r][i] * rows[c][i];
                    MM[r][c] = m;
                }
        }
        if (!chol_solve8<AT>(MM, v)) return false;
    }
    AT gn[8][4];
    for (int j = 0; j < 8; ++j) {
        AT ga[9];
        for (int i = 0; i < 9; ++i) ga[i] = -v[j] * f[i];
        const AT x1 = n[j][0], y1 = n[j][1], x2 = n[j][2], y2 = n[j][3];
        gn[j][0] = ga[0] * x2 + ga[3] * y2 + ga[6];
        gn[j][1] = ga[1] * x2 + ga[4] * y2 + ga[7];
        gn[j][2] = ga[0] * x1 + ga[1] * y1 + ga[2];
        gn[j][3] = ga[3] * x1 + ga[4] * y1 + ga[5];
    }
    // chain through n = r (p - m), r = sqrt2 / mean |p - m|
    AT gc[8][4];
    for (int j = 0; j < 8; ++j) {
        const AT c0 = p[j][0] - h.m[0], c1 = p[j][1] - h.m[1], c2 = p[j][2] - h.m[2], c3 = p[j][3] - h.m[3];
        g_r1 += gn[j][0] * c0 + gn[j][1] * c1;
        g_r2 += gn[j][2] * c2 + gn[j][3] * c3;
        gc[j][0] = h.r1 * gn[j][0]; gc[j][1] = h.r1 * gn[j][1];
        gc[j][2] = h.r2 * gn[j][2]; gc[j][3] = h.r2 * gn[j][3];
    }
    const AT sqrt2 = AT(1.4142135623730951);
    const AT d1 = sqrt2 / h.r1, d2 = sqrt2 / h.r2;
    const AT g_d1 = -(h.r1 / d1) * g_r1, g_d2 = -(h.r2 / d2) * g_r2;
    for (int j = 0; j < 8; ++j) {
        const AT c0 = p[j][0] - h.m[0], c1 = p[j][1] - h.m[1], c2 = p[j][2] - h.m[2], c3 = p[j][3] - h.m[3];
        const AT rho1 = t_sqrt(c0 * c0 + c1 * c1), rho2 = t_sqrt(c2 * c2 + c3 * c3);
        if (rho1 > AT(0)) { gc[j][0] += g_d1 / AT(8) * c0 / rho1; gc[j][1] += g_d1 / AT(8) * c1 / rho1; }
        if (rho2 > AT(0)) { gc[j][2] += g_d2 / AT(8) * c2 / rho2; gc[j][3] += g_d2 / AT(8) * c3 / rho2; }
    }
    for (int c = 0; c < 4; ++c) {
        AT s = AT(0);
        for (int j = 0; j < 8; ++j) s += gc[j][c];
        g_m[c] -= s;
    }
    for (int j = 0; j < 8; ++j)
        for (int c = 0; c < 4; ++c) {
            const AT o = gc[j][c] + g_m[c] / AT(8);
            if (!(o == o)) return false;
            gp[j][c] = T(o);
        }
    return true;
}

// Real roots of c3 x^3 + c2 x^2 + c1 x + c0 (Cardano / trigonometric), ascending.
template <class T>
DRB_HD int cubic_real_roots(T c0, T c1, T c2, T c3, T* x) {
    if (t_abs(c3) < T(1e-30)) return 0;
    const T a = c2 / c3, b = c1 / c3, c = c0 / c3;
    const T q = (a * a - T(3) * b) / T(9);
    const T r = (T(2) * a * a * a - T(9) * a * b + T(27) * c) / T(54);
    const T r2 = r * r, q3 = q * q * q;
    int n;
    if (r2 < q3) {
        const double th = acos((double)(r / t_sqrt(q3)));
        const T sq = T(-2) * t_sqrt(q);
        x[0] = sq * T(cos(th / 3.0)) - a / T(3);
        x[1] = sq * T(cos((th + 6.283185307179586) / 3.0)) - a / T(3);
        x[2] = sq * T(cos((th - 6.283185307179586) / 3.0)) - a / T(3);
        n = 3;
    } else {
        const T A = -(r < T(0) ? T(-1) : T(1)) * T(cbrt((double)(t_abs(r) + t_sqrt(r2 - q3))));
        const T B = (A != T(0)) ? q / A : T(0);
        x[0] = (A + B) - a / T(3);
        n = 1;
    }
    // polish with two Newton steps on the original cubic
    for (int i = 0; i < n; ++i) {
        DRB_UNROLL
        for (int it = 0; it < 2; ++it) {
            const T f = ((c3 * x[i] + c2) * x[i] + c1) * x[i] + c0;
            const T d = (T(3) * c3 * x[i] + T(2) * c2) * x[i] + c1;
            if (t_abs(d) > T(0)) x[i] -= f / d;
        }
    }
    // sort ascending
    if (n == 3) {
        if (x[0] > x[1]) { const T t = x[0]; x[0] = x[1]; x[1] = t; }
        if (x[1] > x[2]) { const T t = x[1]; x[1] = x[2]; x[2] = t; }
        if (x[0] > x[1]) { const T t = x[0]; x[0] = x[1]; x[1] = t; }
    }
    return n;
}

// pts[7][4] -> up to three F (unit norm, det F = 0); returns the count.
template <class T>
DRB_HD int f7_solve(const T (*pts)[4], T (*F)[9]) {
    T n[7][4];
    const HartleyNorm<T> h = hartley_normalize<T, 7>(pts, n);
    T rows[7][9];
    DRB_UNROLL
    for (int j = 0; j < 7; ++j) epipolar_row(n[j][0], n[j][1], n[j][2], n[j][3], rows[j]);
    T N[2][9];
    null_space_rows<T, 7>(rows, N);
    // det(a N0 + (1 - a) N1) = c0 + c1 a + c2 a^2 + c3 a^3, interpolated at a = 0, +-1, +-2
    T dv[5];
    const T as[5] = {T(0), T(1), T(-1), T(2), T(-2)};
    DRB_UNROLL
    for (int s = 0; s < 5; ++s) {
        T M[9];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) M[i] = as[s] * N[0][i] + (T(1) - as[s]) * N[1][i];
        dv[s] = det3(M);
    }
    const T c0 = dv[0];
    const T c2 = T(0.5) * (dv[1] + dv[2]) - dv[0];
    const T c3 = ((dv[3] - dv[4]) * T(0.5) - (dv[1] - dv[2])) / T(6);
    const T c1 = (dv[1] - dv[2]) * T(0.5) - c3;
    T a[3];
    const int nr = cubic_real_roots(c0, c1, c2, c3, a);
    int nout = 0;
    for (int s = 0; s < nr; ++s) {
        T Fn[9], Fd[9];
        T n2 = T(0);
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) Fn[i] = a[s] * N[0][i] + (T(1) - a[s]) * N[1][i];
        denormalize_f(Fn, h, Fd);
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) n2 += Fd[i] * Fd[i];
        if (!(n2 > T(0)) || !(n2 < T(1e37))) continue;
        const T inv = t_rsqrt(n2);
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) F[nout][i] = Fd[i] * inv;
        ++nout;
    }
    for (int s = nout; s < 3; ++s) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) F[s][i] = (i % 4 == 0) ? T(1) : T(0);
    }
    return nout;
}

}  // namespace drb
