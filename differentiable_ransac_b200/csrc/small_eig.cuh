// Eigen-decomposition of small symmetric matrices by cyclic Jacobi rotations (one thread, any scalar type).
// Used where the reference calls torch.linalg.svd / cv2 SVD on tiny matrices OUTSIDE the hot loop: the 9 x 9
// moment matrix of the non-minimal fits (refit_math.cuh), E^T E and the 4 x 4 DLT systems of pose recovery
// (pose_math.cuh).  Jacobi computes small eigenvalues of positive semi-definite matrices to high RELATIVE
// accuracy, which is what a null-space extraction needs.
#pragma once

#include "drb_common.cuh"

namespace drb {

// `A` (row-major N x N, symmetric) is overwritten by its diagonal form; `V` (row-major) receives the
// eigenvectors as COLUMNS.  Stops when the off-diagonal mass is below rounding, or after max_sweeps.
template <class T, int N>
DRB_HD void jacobi_eig(T* A, T* V, int max_sweeps = 16) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? T(1) : T(0);
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        T off = T(0), diag = T(0);
        for (int i = 0; i < N; ++i) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
        }
        if (!(off > diag * T(1e-34))) break;
        for (int p = 0; p < N - 1; ++p) {
            for (int q = p + 1; q < N; ++q) {
                const T apq = A[p * N + q];
                if (apq == T(0)) continue;
                const T theta = (A[q * N + q] - A[p * N + p]) / (T(2) * apq);
                const T t = (theta >= T(0) ? T(1) : T(-1)) / (t_abs(theta) + t_sqrt(theta * theta + T(1)));
                const T c = T(1) / t_sqrt(t * t + T(1));
                const T s = t * c;
                for (int k = 0; k < N; ++k) {
                    const T akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const T apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const T vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
        }
    }
}

}  // namespace drb
