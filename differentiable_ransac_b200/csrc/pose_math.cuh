// Relative pose from an essential matrix: decomposition, DLT triangulation, cheirality vote, angular
// errors.  SURVEY 8f ranks 2-3: replaces
//   * cv_utils.recoverPose / decompose_E / cheirality_check (cv_utils.py:48-116, :179-189), which calls
//     cv2.triangulatePoints per candidate pose on the host,
//   * cv2.recoverPose as used for the ground-truth inlier mask of MatchLoss (loss.py:126-135),
//   * evaluate_R_t_tensor (cv_utils.py:361-378) behind eval_essential_matrix (cv_utils.py:503-525).
// One thread runs any of these functions; pose.cu spreads the correspondences over a CTA.
#pragma once

#include "drb_common.cuh"
#include "dual.cuh"
#include "small_eig.cuh"

namespace drb {

DRB_HD float t_acos(float x) { return acosf(x); }
DRB_HD double t_acos(double x) { return acos(x); }
template <class T, int P>
DRB_HD Dual<T, P> t_acos(const Dual<T, P>& a) {
    Dual<T, P> r;
    r.v = t_acos(a.v);
    const T g = T(-1) / t_sqrt(T(1) - a.v * a.v);   // infinite at |x| = 1, as torch.arccos's backward
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * g;
    return r;
}

template <class T>
struct PoseCandidates {
    T R1[9], R2[9], t[3];  // the four poses are (R1,t), (R2,t), (R1,-t), (R2,-t), in the reference's order
};

template <class T>
DRB_HD void cross3(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// E = U diag(s) V^T with det U = det V = +1;  R1 = U W V^T, R2 = U W^T V^T, t = U[:,2]
// (cv_utils.py:83-116; W = [[0,-1,0],[1,0,0],[0,0,1]]).  U and V come from the eigenvectors of E^T E, so the
// pair {R1, R2} and the line of t are those of any SVD; which of the two is called R1 and the sign of t are as
// arbitrary as LAPACK's -- the cheirality vote over the four combinations removes both.
template <class T>
DRB_HD bool decompose_essential(const T* E, PoseCandidates<T>& pc) {
    T A[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = T(0);
            for (int k = 0; k < 3; ++k) s += E[k * 3 + i] * E[k * 3 + j];
            A[i * 3 + j] = s;
        }
    jacobi_eig<T, 3>(A, V);
    int o0 = 0, o1 = 1, o2 = 2;  // descending eigenvalues
    if (A[o0 * 4] < A[o1 * 4]) { const int t = o0; o0 = o1; o1 = t; }
    if (A[o1 * 4] < A[o2 * 4]) { const int t = o1; o1 = o2; o2 = t; }
    if (A[o0 * 4] < A[o1 * 4]) { const int t = o0; o0 = o1; o1 = t; }
    T v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
    for (int k = 0; k < 3; ++k) { v1[k] = V[k * 3 + o0]; v2[k] = V[k * 3 + o1]; }
    cross3(v1, v2, v3);
    T n1 = T(0), n2 = T(0), d = T(0);
    for (int i = 0; i < 3; ++i) {
        u1[i] = E[i * 3] * v1[0] + E[i * 3 + 1] * v1[1] + E[i * 3 + 2] * v1[2];
        u2[i] = E[i * 3] * v2[0] + E[i * 3 + 1] * v2[1] + E[i * 3 + 2] * v2[2];
        n1 += u1[i] * u1[i];
    }
    if (!(n1 > T(0))) return false;
    n1 = t_sqrt(n1);
    for (int i = 0; i < 3; ++i) { u1[i] /= n1; d += u1[i] * u2[i]; }
    for (int i = 0; i < 3; ++i) { u2[i] -= d * u1[i]; n2 += u2[i] * u2[i]; }
    if (!(n2 > T(0))) return false;
    n2 = t_sqrt(n2);
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
    cross3(u1, u2, u3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const T a = u2[i] * v1[j] - u1[i] * v2[j], b = u3[i] * v3[j];
            pc.R1[i * 3 + j] = a + b;
            pc.R2[i * 3 + j] = b - a;
        }
    bool ok = true;
    for (int i = 0; i < 3; ++i) pc.t[i] = u3[i];
    for (int i = 0; i < 9; ++i) ok = ok && (pc.R1[i] == pc.R1[i]) && (pc.R2[i] == pc.R2[i]);
    return ok;
}

// Horn's closed form (cv_utils.new_decompose_E, cv_utils.py:118-165; what PoseLoss uses, loss.py:17-30 `svd=False`):
// b = sqrt(tr(E E^T)/2) * (e_i x e_j)/|e_i x e_j| for the largest pairwise cross product of the COLUMNS of E,
// R1,2 = (Cof(E) -+ [b]x E) / (b.b), t = b/|b|.  The reference builds [b]x with torch.tensor(...) of tensor
// elements, which cuts the graph: [b]x is a constant for the gradient (detach_value), everything else is
// differentiated.  Cof(E) is formed from cross products of the rows (the reference: inv(E)^T det(E), which is
// the same matrix computed through a nearly singular inverse).
template <class T>
DRB_HD bool decompose_essential_horn(const T* E, PoseCandidates<T>& pc) {
    const T e1[3] = {E[0], E[3], E[6]}, e2[3] = {E[1], E[4], E[7]}, e3[3] = {E[2], E[5], E[8]};
    T c[3][3];
    cross3(e1, e2, c[0]);
    cross3(e2, e3, c[1]);
    cross3(e3, e1, c[2]);
    T nrm[3];
    for (int k = 0; k < 3; ++k) nrm[k] = t_sqrt(c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2]);
    int big = 0;  // torch.argmax: first maximum
    if (nrm[1] > nrm[big]) big = 1;
    if (nrm[2] > nrm[big]) big = 2;
    if (!(nrm[big] > T(0))) return false;
    T tr = T(0);
    for (int i = 0; i < 9; ++i) tr += E[i] * E[i];
    const T scale = t_sqrt(T(0.5) * tr);
    T b[3], bb = T(0);
    for (int i = 0; i < 3; ++i) {
        b[i] = scale * c[big][i] / nrm[big];
        bb += b[i] * b[i];
    }
    const T nb = t_sqrt(bb);
    for (int i = 0; i < 3; ++i) pc.t[i] = b[i] / nb;
    const T b0 = detach_value(b[0]), b1 = detach_value(b[1]), b2 = detach_value(b[2]);
    const T Bx[9] = {T(0), -b2, b1, b2, T(0), -b0, -b1, b0, T(0)};
    T cof[9];
    cross3(E + 3, E + 6, cof);      // rows of the cofactor matrix = cross products of the other two rows
    cross3(E + 6, E, cof + 3);
    cross3(E, E + 3, cof + 6);
    bool ok = true;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const T be = Bx[i * 3] * E[j] + Bx[i * 3 + 1] * E[3 + j] + Bx[i * 3 + 2] * E[6 + j];
            pc.R1[i * 3 + j] = (cof[i * 3 + j] - be) / bb;
            pc.R2[i * 3 + j] = (cof[i * 3 + j] + be) / bb;
            ok = ok && (pc.R1[i * 3 + j] == pc.R1[i * 3 + j]) && (pc.R2[i * 3 + j] == pc.R2[i * 3 + j]);
        }
    return ok;
}

// Homogeneous DLT triangulation between P0 = [I | 0] and P = [R | t] (what cv2.triangulatePoints computes,
// cv_utils.py:182): the right singular vector of the smallest singular value of the 4 x 4 system
// (x P[2] - P[0]; y P[2] - P[1]) of both views.  Sign of Q arbitrary.
template <class T>
DRB_HD void triangulate_dlt(const T* R, const T* t, T x1, T y1, T x2, T y2, T* Q) {
    T rows[4][4] = {{T(-1), T(0), x1, T(0)},
                    {T(0), T(-1), y1, T(0)},
                    {x2 * R[6] - R[0], x2 * R[7] - R[1], x2 * R[8] - R[2], x2 * t[2] - t[0]},
                    {y2 * R[6] - R[3], y2 * R[7] - R[4], y2 * R[8] - R[5], y2 * t[2] - t[1]}};
    T A[16], V[16];
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            T s = T(0);
            for (int k = 0; k < 4; ++k) s += rows[k][i] * rows[k][j];
            A[i * 4 + j] = A[j * 4 + i] = s;
        }
    jacobi_eig<T, 4>(A, V);
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (A[i * 5] < A[best * 5]) best = i;
    for (int k = 0; k < 4; ++k) Q[k] = V[k * 4 + best];
}

// The correspondence lies in front of both cameras, closer than `dist`, under pose c of the four
// (cv_utils.py:184-187: Q_z Q_w > 0, Q_z / Q_w < dist, 0 < (P Q)_z < dist).
template <class T>
DRB_HD bool cheirality_one(const PoseCandidates<T>& pc, int c, T x1, T y1, T x2, T y2, T dist) {
    const T* R = (c & 1) ? pc.R2 : pc.R1;
    const T sg = (c & 2) ? T(-1) : T(1);
    const T t[3] = {sg * pc.t[0], sg * pc.t[1], sg * pc.t[2]};
    T Q[4];
    triangulate_dlt<T>(R, t, x1, y1, x2, y2, Q);
    const T X = Q[0] / Q[3], Y = Q[1] / Q[3], Z = Q[2] / Q[3];
    const T z2 = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    return Q[2] * Q[3] > T(0) && Z < dist && z2 > T(0) && z2 < dist;
}

// Bit c of the result = cheirality_one under pose c.
template <class T>
DRB_HD int cheirality_bits(const PoseCandidates<T>& pc, T x1, T y1, T x2, T y2, T dist) {
    int bits = 0;
    for (int c = 0; c < 4; ++c)
        if (cheirality_one<T>(pc, c, x1, y1, x2, y2, dist)) bits |= 1 << c;
    return bits;
}

// Rotation / translation angular errors in DEGREES, evaluate_R_t_tensor (cv_utils.py:361-378) followed by the
// conversion of eval_essential_matrix (cv_utils.py:525).
template <class T>
DRB_HD void pose_errors_deg(const T* R, const T* t, const T* R_gt, const T* t_gt, T& err_r, T& err_t) {
    const T eps = T(1e-8), deg = T(57.29577951308232);
    T tr = T(0), dot = T(0), ng = T(0);
    for (int i = 0; i < 9; ++i) tr += R[i] * R_gt[i];  // trace(R R_gt^T)
    T c = (tr - T(1)) * T(0.5);
    c = c > T(1) ? T(1) : (c < T(-1) ? T(-1) : c);
    err_r = t_acos(c) * deg;
    for (int i = 0; i < 3; ++i) ng += t_gt[i] * t_gt[i];
    ng = t_sqrt(ng) + eps;
    for (int i = 0; i < 3; ++i) dot += t[i] * (t_gt[i] / ng);
    T loss = T(1) - dot * dot;
    loss = loss > eps ? loss : eps;
    err_t = t_acos(t_sqrt(T(1) - loss + eps)) * deg;
}

}  // namespace drb
