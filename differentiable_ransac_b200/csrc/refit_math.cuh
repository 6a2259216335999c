// Non-minimal model fits on a SET of correspondences (the winner's inliers, or all points): the refit
// and local-optimisation steps that follow the hypothesize-and-score loop in test mode.
//
// Replaces, for n > sample_size points,
//   * EssentialMatrixEstimatorNister.estimate_model without pymagsac
//     (estimators/essential_matrix_estimator_nister.py:51-65 -> :69-430 on n rows: the last four right
//     singular vectors of A^T A span the solution space, then the same polynomial system as the 5-point),
//   * FundamentalMatrixEstimatorNew.normalize + estimate_non_minimal_model
//     (estimators/fundamental_matrix_estimator.py:177-260: Hartley normalisation over the n points, last
//     right singular vector of A^T A, F = T2^T F T1),
// as called from ransac.py:148-165 (final refit) and ransac.py:217-257 (lo = 1, 2).
//
// The 9 x 9 moment matrix A^T A is accumulated by the whole CTA (refit.cu); everything here is the serial
// tail one thread runs per pair, in double: cyclic Jacobi for the eigenvectors, then the shared five-point
// machinery of e5_math.cuh.
#pragma once

#include "drb_common.cuh"
#include "e5_math.cuh"
#include "f8_math.cuh"
#include "small_eig.cuh"

namespace drb {

// Packed index of the symmetric 9 x 9 moment matrix, i <= j.
DRB_HD int sym9(int i, int j) { return i * 9 - (i * (i - 1)) / 2 + (j - i); }

// The `count` eigenvectors of the packed moment matrix with the smallest eigenvalues, written in the
// reference's order (`v[:, -count:, :]` of torch.linalg.svd: descending singular values, so out[count-1] is
// the smallest).  Sign of each vector arbitrary, as LAPACK's is.
template <class T>
DRB_HD void smallest_eigenvectors9(const T* packed, int count, T (*out)[9]) {
    T A[81], V[81];
    for (int i = 0; i < 9; ++i)
        for (int j = i; j < 9; ++j) A[i * 9 + j] = A[j * 9 + i] = packed[sym9(i, j)];
    jacobi_eig<T, 9>(A, V);
    bool used[9];
    for (int i = 0; i < 9; ++i) used[i] = false;
    for (int r = 0; r < count; ++r) {
        int best = -1;
        for (int i = 0; i < 9; ++i)
            if (!used[i] && (best < 0 || A[i * 9 + i] < A[best * 9 + best])) best = i;
        used[best] = true;
        for (int k = 0; k < 9; ++k) out[count - 1 - r][k] = V[k * 9 + best];
    }
}

template <class T>
struct LocalMat {
    T a[200];
    DRB_HD T& operator()(int r, int c) { return a[r * 20 + c]; }
};

// Essential matrices consistent with the moment matrix of n >= 5 correspondences: up to 10 unit-norm
// models (x2^T E x1 = 0) in `models`, return value = how many.  Slots beyond are not written.
template <class T>
DRB_HD int e5_refit_from_moments(const T* packed, T (*models)[9], int polish_iters = 2) {
    E5Sample<T> S;
    smallest_eigenvectors9<T>(packed, 4, S.N);
    LocalMat<T> M;
    if (!e5_prepare_from_null<T, LocalMat<T>>(M, S)) return 0;
    T blo[10], bhi[10];
    int n0 = 0;
    const int nb = isolate_deg10<T>(S.P, blo, bhi, n0);
    int nout = 0;
    for (int r = 0; r < nb; ++r) {
        T zr;
        if (!root_from_bracket<T>(S.P, r >= n0, blo[r], bhi[r], zr)) continue;
        if (!e5_model_from_root<T>(S, zr, polish_iters, models[nout])) continue;
        ++nout;
    }
    return nout;
}

// Fundamental matrix from the moment matrix of Hartley-normalised correspondences + the normalisation.
template <class T>
DRB_HD bool f8_refit_from_moments(const T* packed, const HartleyNorm<T>& h, T* F) {
    T f[1][9];
    smallest_eigenvectors9<T>(packed, 1, f);
    denormalize_f(f[0], h, F);
    bool ok = true;
    for (int i = 0; i < 9; ++i) ok = ok && (F[i] == F[i]) && (t_abs(F[i]) < T(1e30));
    return ok;
}

}  // namespace drb
