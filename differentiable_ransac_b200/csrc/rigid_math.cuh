// Three-point rigid transformation (forward + backward), one hypothesis per thread.
//
// Replaces RigidTransformationSVDBasedSolver.estimate_model
// (estimators/rigid_transformation_SVD_based_solver.py:11-74), BOTH `flag` branches:
//   flag != 0 (the default, the one RANSAC3D uses, ransac.py:367): the reference takes the
//     SVD of cov^T cov, a symmetric PSD matrix, so R = V U^T is the identity up to rounding
//      and the model is the centroid translation; emitted here as exactly I.
//   flag == 0: SVD of cov^T = U S V^T, R = V U^T with the reflection fix of :59-62, i.e. the
//     orthogonal polar factor of cov with det +1.  Computed from the symmetric eigen
//     decomposition of cov cov^T (cyclic Jacobi) -- the third singular pair is rebuilt by
//     cross products, which is exactly what the det-fix selects.
// The translation keeps the reference's broadcast (:66): t_j = c1_j - c0_j * sum_i R_ij.
#pragma once

#include "drb_common.cuh"

namespace drb {

// Symmetric 3x3 eigen decomposition by cyclic Jacobi.  A row-major (symmetric), on exit
// w = eigenvalues, V columns = eigenvectors (row-major V[3*i+j] = component i of vector j).
template <class T>
DRB_HD void jacobi_eig3(const T* Ain, T* w, T* V) {
    T a[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) { a[i] = Ain[i]; V[i] = (i % 4 == 0) ? T(1) : T(0); }
    for (int sweep = 0; sweep < 8; ++sweep) {
        DRB_UNROLL
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            const T apq = a[3 * p + q];
            if (t_abs(apq) < T(1e-30)) continue;
            const T theta = (a[3 * q + q] - a[3 * p + p]) / (T(2) * apq);
            const T t = (theta >= T(0) ? T(1) : T(-1)) / (t_abs(theta) + t_sqrt(theta * theta + T(1)));
            const T c = T(1) / t_sqrt(t * t + T(1));
            const T s = t * c;
            DRB_UNROLL
            for (int k = 0; k < 3; ++k) {  // A <- A J
                const T akp = a[3 * k + p], akq = a[3 * k + q];
                a[3 * k + p] = c * akp - s * akq;
                a[3 * k + q] = s * akp + c * akq;
            }
            DRB_UNROLL
            for (int k = 0; k < 3; ++k) {  // A <- J^T A
                const T apk = a[3 * p + k], aqk = a[3 * q + k];
                a[3 * p + k] = c * apk - s * aqk;
                a[3 * q + k] = s * apk + c * aqk;
            }
            DRB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const T vkp = V[3 * k + p], vkq = V[3 * k + q];
                V[3 * k + p] = c * vkp - s * vkq;
                V[3 * k + q] = s * vkp + c * vkq;
            }
        }
    }
    w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
}

template <class T>
struct RigidPolar {
    T Q[9];     // rotation: orthogonal polar factor of cov (cov = Q P), det +1
    T Ur[9];    // right singular vectors of cov as columns (eigenvectors of P)
    T sig[3];   // signed singular values (third one carries the det-fix sign)
    bool ok;
};

// cov (row-major 3x3) -> polar rotation with det +1 and the pieces the backward needs.
template <class T>
DRB_HD RigidPolar<T> polar_rotation(const T* cov) {
    RigidPolar<T> r;
    T C[9], w[3], V[9];
    mul33_nt(cov, cov, C);  // cov cov^T = V S^2 V^T  (V = left singular vectors of cov)
    jacobi_eig3(C, w, V);
    // order eigenvalues descending
    int o0 = 0, o1 = 1, o2 = 2;
    if (w[o0] < w[o1]) { const int t = o0; o0 = o1; o1 = t; }
    if (w[o1] < w[o2]) { const int t = o1; o1 = o2; o2 = t; }
    if (w[o0] < w[o1]) { const int t = o0; o0 = o1; o1 = t; }
    T v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        v1[i] = o0 == 0 ? V[3 * i] : (o0 == 1 ? V[3 * i + 1] : V[3 * i + 2]);
        v2[i] = o1 == 0 ? V[3 * i] : (o1 == 1 ? V[3 * i + 1] : V[3 * i + 2]);
    }
    const T l1 = o0 == 0 ? w[0] : (o0 == 1 ? w[1] : w[2]);
    const T l2 = o1 == 0 ? w[0] : (o1 == 1 ? w[1] : w[2]);
    const T l3 = o2 == 0 ? w[0] : (o2 == 1 ? w[1] : w[2]);
    const T s1 = t_sqrt(t_max(l1, T(0))), s2 = t_sqrt(t_max(l2, T(0)));
    r.ok = (s1 > T(0)) && (s2 > T(1e-6) * s1);
    // u_i = cov^T v_i / s_i
    T n1 = T(0), n2 = T(0), d12 = T(0);
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        u1[i] = cov[i] * v1[0] + cov[3 + i] * v1[1] + cov[6 + i] * v1[2];
        u2[i] = cov[i] * v2[0] + cov[3 + i] * v2[1] + cov[6 + i] * v2[2];
        n1 += u1[i] * u1[i];
    }
    const T in1 = n1 > T(0) ? T(1) / t_sqrt(n1) : T(0);
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) { u1[i] *= in1; d12 += u1[i] * u2[i]; }
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) { u2[i] -= d12 * u1[i]; n2 += u2[i] * u2[i]; }
    const T in2 = n2 > T(0) ? T(1) / t_sqrt(n2) : T(0);
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) u2[i] *= in2;
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    // signed third singular value: u3^T cov^T v3
    T s3 = T(0);
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) s3 += u3[i] * (cov[i] * v3[0] + cov[3 + i] * v3[1] + cov[6 + i] * v3[2]);
    (void)l3;
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) r.Q[3 * i + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
        r.Ur[3 * i] = u1[i]; r.Ur[3 * i + 1] = u2[i]; r.Ur[3 * i + 2] = u3[i];
    }
    r.sig[0] = s1; r.sig[1] = s2; r.sig[2] = s3;
    return r;
}

template <class T>
struct RigidStats {
    T c[6];      // centroid
    T rho0, rho1;
    T cov[9];
};

template <class T, int S>
DRB_HD RigidStats<T> rigid_stats(const T (*pts)[6]) {
    RigidStats<T> st;
    DRB_UNROLL
    for (int c = 0; c < 6; ++c) {
        T s = T(0);
        DRB_UNROLL
        for (int j = 0; j < S; ++j) s += pts[j][c];
        st.c[c] = s / T(S);
    }
    T a0 = T(0), a1 = T(0);
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) st.cov[i] = T(0);
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        T d[6];
        DRB_UNROLL
        for (int c = 0; c < 6; ++c) d[c] = pts[j][c] - st.c[c];
        a0 += t_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        a1 += t_sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
        DRB_UNROLL
        for (int i = 0; i < 3; ++i) {
            DRB_UNROLL
            for (int k = 0; k < 3; ++k) st.cov[3 * i + k] += d[i] * d[3 + k];
        }
    }
    const T sqrt3 = T(1.7320508075688772);
    st.rho0 = sqrt3 / (a0 / T(S));
    st.rho1 = sqrt3 / (a1 / T(S));
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) st.cov[i] *= st.rho0 * st.rho1;
    return st;
}

// pts[3][6] = [P | Q] -> 4x4 pose row-major.  Returns validity (finite and, for flag == 0,
// a rank-2 covariance).
template <class T>
DRB_HD bool rigid3_solve(const T (*pts)[6], int flag, T* model) {
    const RigidStats<T> st = rigid_stats<T, 3>(pts);
    T R[9];
    bool ok = true;
    if (flag) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? T(1) : T(0);
    } else {
        // cov cov^T squares the condition number: run the 3x3 eigen problem in double
        double covd[9];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) covd[i] = (double)st.cov[i];
        const RigidPolar<double> pr = polar_rotation<double>(covd);
        ok = pr.ok;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) R[i] = T(pr.Q[i]);
    }
    DRB_UNROLL
    for (int j = 0; j < 3; ++j) {
        const T colsum = R[j] + R[3 + j] + R[6 + j];
        const T t = st.c[3 + j] - st.c[j] * colsum;
        model[j] = R[j]; model[4 + j] = R[3 + j]; model[8 + j] = R[6 + j];
        model[4 * j + 3] = t;
    }
    model[12] = T(0); model[13] = T(0); model[14] = T(0); model[15] = T(1);
    DRB_UNROLL
    for (int i = 0; i < 12; ++i) ok = ok && (model[i] == model[i]) && (t_abs(model[i]) < T(1e30));
    // the reference drops samples whose covariance is NaN (:45); degenerate distances do that
    ok = ok && (st.rho0 == st.rho0) && (st.rho1 == st.rho1) && (t_abs(st.rho0) < T(1e30)) && (t_abs(st.rho1) < T(1e30));
    return ok;
}

// g = dL/dmodel (4x4 row-major, only the [R|t] block is read) -> gp[3][6].
template <class T, class AT = double>
DRB_HD bool rigid3_backward(const T (*pts)[6], int flag, const T* g, T (*gp)[6]) {
    AT p[3][6];
    for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 6; ++c) p[j][c] = AT(pts[j][c]);
    const RigidStats<AT> st = rigid_stats<AT, 3>(p);
    AT gR[9], gt[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) gR[3 * i + j] = AT(g[4 * i + j]);
        gt[i] = AT(g[4 * i + 3]);
    }
    AT R[9];
    RigidPolar<AT> pr;
    if (flag) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? AT(1) : AT(0);
    } else {
        pr = polar_rotation<AT>(st.cov);
        if (!pr.ok) return false;
        for (int i = 0; i < 9; ++i) R[i] = pr.Q[i];
    }
    AT gc[6];
    for (int j = 0; j < 3; ++j) {
        const AT colsum = R[j] + R[3 + j] + R[6 + j];
        gc[3 + j] = gt[j];
        gc[j] = -gt[j] * colsum;
        for (int i = 0; i < 3; ++i) gR[3 * i + j] += -gt[j] * st.c[j];
    }
    AT gd[3][6];
    for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 6; ++c) gd[j][c] = AT(0);
    if (!flag) {
        // dL/dcov = 2 Q W,  W = U Wt U^T,  Wt_ij = (U^T skew(Q^T gQ) U)_ij / (s_i + s_j)
        AT QtG[9], Z[9], tmp[9], Zt[9], W[9], gcov[9];
        mul33_tn(pr.Q, gR, QtG);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Z[3 * i + j] = AT(0.5) * (QtG[3 * i + j] - QtG[3 * j + i]);
        mul33_tn(pr.Ur, Z, tmp);
        mul33(tmp, pr.Ur, Zt);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const AT den = pr.sig[i] + pr.sig[j];
                Zt[3 * i + j] = (i != j && t_abs(den) > AT(0)) ? Zt[3 * i + j] / den : AT(0);
            }
        mul33(pr.Ur, Zt, tmp);
        mul33_nt(tmp, pr.Ur, W);
        mul33(pr.Q, W, gcov);
        for (int i = 0; i < 9; ++i) gcov[i] *= AT(2);
        // cov = rho0 rho1 sum_j a_j b_j^T ; R is invariant to the positive scale rho0 rho1
        const AT sc = st.rho0 * st.rho1;
        for (int j = 0; j < 3; ++j) {
            AT a[3], b[3];
            for (int i = 0; i < 3; ++i) { a[i] = p[j][i] - st.c[i]; b[i] = p[j][3 + i] - st.c[3 + i]; }
            for (int i = 0; i < 3; ++i) {
                gd[j][i] = sc * (gcov[3 * i] * b[0] + gcov[3 * i + 1] * b[1] + gcov[3 * i + 2] * b[2]);
                gd[j][3 + i] = sc * (gcov[i] * a[0] + gcov[3 + i] * a[1] + gcov[6 + i] * a[2]);
            }
        }
    }
    // d_j = p_j - c, c = mean p  =>  dL/dp_j = gd_j - mean(gd) + gc / 3
    for (int c = 0; c < 6; ++c) {
        const AT mean = (gd[0][c] + gd[1][c] + gd[2][c]) / AT(3);
        for (int j = 0; j < 3; ++j) {
            const AT o = gd[j][c] - mean + gc[c] / AT(3);
            if (!(o == o)) return false;
            gp[j][c] = T(o);
        }
    }
    return true;
}

// ---- rigid residual from the points' second moments (score.cu: rigid_residual_moments_kernel) -------------------
// sum_n ||q_n - (R p_n + t)||^2 and its gradient in (R, t) are quadratic forms in the model whose coefficients are 22
// sums over the points alone (rigid_transformation_SVD_based_solver.py:76-89 computes the same sum point by point):
//     mom[0..5] = S_pp (xx, xy, xz, yy, yz, zz),  mom[6..14] = S_qp row-major,  mom[15..17] = s_p,  mom[18..20] = s_q,
//     mom[21] = sum |q|^2
//     sum d_i p_j = S_qp[i][j] - R_i . S_pp[:, j] - t_i s_p[j],      sum d_i = s_q[i] - R_i . s_p - N t_i
//     sum |d|^2   = s_qq - 2 sum_i (R_i . S_qp[i] + t_i s_q[i]) + sum_i (R_i S_pp R_i' + 2 t_i R_i . s_p + N t_i^2)
// AT is the accumulation type: double on the device (the sums cancel for a model that fits).
constexpr int kRigidMoments = 22;

template <class AT>
DRB_HD void rigid_moments_add(const AT* p, const AT* q, AT* s) {
    s[0] += p[0] * p[0]; s[1] += p[0] * p[1]; s[2] += p[0] * p[2];
    s[3] += p[1] * p[1]; s[4] += p[1] * p[2]; s[5] += p[2] * p[2];
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) s[6 + 3 * i + j] += q[i] * p[j];
        s[15 + i] += p[i];
        s[18 + i] += q[i];
        s[21] += q[i] * q[i];
    }
}

// model12: the 3 x 4 block [R | t] row-major (the first twelve entries of the 4 x 4 model).  res = sum |d|^2;
// g12[4 i + j] = -sum d_i p_j (j < 3), g12[4 i + 3] = -sum d_i: half of d res / d model12.
template <class T, class AT>
DRB_HD void rigid_residual_from_moments(const AT* mom, AT n_points, const T* model12, AT& res, AT* g12) {
    const AT Spp[3][3] = {{mom[0], mom[1], mom[2]}, {mom[1], mom[3], mom[4]}, {mom[2], mom[4], mom[5]}};
    res = mom[21];
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        const AT r[3] = {AT(model12[4 * i]), AT(model12[4 * i + 1]), AT(model12[4 * i + 2])};
        const AT t = AT(model12[4 * i + 3]);
        AT rS[3];                    // R_i . S_pp[:, j]
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) rS[j] = r[0] * Spp[0][j] + r[1] * Spp[1][j] + r[2] * Spp[2][j];
        const AT r_sp = r[0] * mom[15] + r[1] * mom[16] + r[2] * mom[17];
        const AT r_sqp = r[0] * mom[6 + 3 * i] + r[1] * mom[7 + 3 * i] + r[2] * mom[8 + 3 * i];
        res += AT(-2) * (r_sqp + t * mom[18 + i]) + (rS[0] * r[0] + rS[1] * r[1] + rS[2] * r[2]) + AT(2) * t * r_sp +
               n_points * t * t;
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) g12[4 * i + j] = -(mom[6 + 3 * i + j] - rS[j] - t * mom[15 + j]);
        g12[4 * i + 3] = -(mom[18 + i] - r_sp - n_points * t);
    }
}

}  // namespace drb
