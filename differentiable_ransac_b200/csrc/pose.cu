// Pose recovery on the device: one CTA per (pair, model).  Thread 0 decomposes the
// essential matrix; the CTA's threads triangulate every correspondence under the four candidate poses
// (DLT, double) and vote; the winning pose, its cheirality mask and -- given a ground-truth pose -- the
// angular errors are written out.  See pose_math.cuh for the reference lines this replaces.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "pose_math.cuh"

namespace drb {

constexpr int kPoseThreads = 128;

// HORN selects the decomposition: false = SVD-equivalent (cv_utils.decompose_E), true = Horn's closed form
// (cv_utils.new_decompose_E, what PoseLoss differentiates).  With `grad` (HORN only) thread 0 re-runs the
// closed form on forward-mode duals and writes d((err_R + err_t) / 2)/dE: the backward of one PoseLoss term.
template <bool HORN>
__global__ void __launch_bounds__(kPoseThreads)
recover_pose_kernel(const float* __restrict__ E, const float* __restrict__ matches, const int32_t* __restrict__ npts,
                    const float* __restrict__ R_gt, const float* __restrict__ t_gt, int M, int N, float dist,
                    float* __restrict__ R, float* __restrict__ t, uint8_t* __restrict__ mask,
                    int32_t* __restrict__ ngood, float* __restrict__ err, float* __restrict__ grad) {
    __shared__ PoseCandidates<double> pc;
    __shared__ int ok_s;
    __shared__ int warp_cnt[kPoseThreads / 32][4];
    __shared__ int best_s;
    const int m = blockIdx.x, b = blockIdx.y;
    const size_t bm = (size_t)b * M + m;
    if (threadIdx.x == 0) {
        double e[9];
        for (int i = 0; i < 9; ++i) e[i] = (double)E[bm * 9 + i];
        bool fin = true;
        for (int i = 0; i < 9; ++i) fin = fin && (e[i] == e[i]) && (fabs(e[i]) < 1e30);
        ok_s = (fin && (HORN ? decompose_essential_horn<double>(e, pc) : decompose_essential<double>(e, pc))) ? 1 : 0;
    }
    __syncthreads();
    const int n_used = npts ? min(npts[b], N) : N;
    uint8_t* mk = mask ? mask + bm * N : nullptr;
    if (!ok_s) {  // not decomposable: identity pose, nothing in front of anything
        if (threadIdx.x == 0) {
            if (R)
                for (int i = 0; i < 9; ++i) R[bm * 9 + i] = (i % 4 == 0) ? 1.f : 0.f;
            if (t)
                for (int i = 0; i < 3; ++i) t[bm * 3 + i] = 0.f;
            if (ngood) ngood[bm] = 0;
            if (err) { err[bm * 2] = 180.f; err[bm * 2 + 1] = 90.f; }  // eval_essential_matrix's failure values
            if (grad)
                for (int i = 0; i < 9; ++i) grad[bm * 9 + i] = 0.f;
        }
        if (mk)
            for (int n = threadIdx.x; n < N; n += kPoseThreads) mk[n] = 0;
        return;
    }
    const float4* pts = reinterpret_cast<const float4*>(matches) + (size_t)b * N;
    int cnt[4] = {0, 0, 0, 0};
    for (int n = threadIdx.x; n < N; n += kPoseThreads) {
        int bits = 0;
        if (n < n_used) {
            const float4 p = __ldg(pts + n);
            bits = cheirality_bits<double>(pc, (double)p.x, (double)p.y, (double)p.z, (double)p.w, (double)dist);
        }
        if (mk) mk[n] = (uint8_t)bits;
        DRB_UNROLL
        for (int c = 0; c < 4; ++c) cnt[c] += (bits >> c) & 1;
    }
    DRB_UNROLL
    for (int c = 0; c < 4; ++c) {
        DRB_UNROLL
        for (int o = 16; o > 0; o >>= 1) cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
        if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5][c] = cnt[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot[4], best = 0;
        for (int c = 0; c < 4; ++c) {
            tot[c] = 0;
            for (int w = 0; w < kPoseThreads / 32; ++w) tot[c] += warp_cnt[w][c];
        }
        for (int c = 1; c < 4; ++c)
            if (tot[c] > tot[best]) best = c;  // first maximum, like torch.argmax (cv_utils.py:71)
        best_s = best;
        double Rb[9], tb[3];
        for (int i = 0; i < 9; ++i) Rb[i] = (best & 1) ? pc.R2[i] : pc.R1[i];
        for (int i = 0; i < 3; ++i) tb[i] = (best & 2) ? -pc.t[i] : pc.t[i];
        if (R)
            for (int i = 0; i < 9; ++i) R[bm * 9 + i] = (float)Rb[i];
        if (t)
            for (int i = 0; i < 3; ++i) t[bm * 3 + i] = (float)tb[i];
        if (ngood) ngood[bm] = tot[best];
        if (err && R_gt && t_gt) {
            double Rg[9], tg[3], er, et;
            for (int i = 0; i < 9; ++i) Rg[i] = (double)R_gt[b * 9 + i];
            for (int i = 0; i < 3; ++i) tg[i] = (double)t_gt[b * 3 + i];
            pose_errors_deg<double>(Rb, tb, Rg, tg, er, et);
            err[bm * 2] = (float)er;
            err[bm * 2 + 1] = (float)et;
            if (HORN && grad) {
                typedef Dual<double, 9> D;
                D Ed[9], Rgd[9], tgd[3], Rd[9], td[3], erd, etd;
                for (int i = 0; i < 9; ++i) { Ed[i] = D::variable((double)E[bm * 9 + i], i); Rgd[i] = D(Rg[i]); }
                for (int i = 0; i < 3; ++i) tgd[i] = D(tg[i]);
                PoseCandidates<D> pd;
                decompose_essential_horn<D>(Ed, pd);
                for (int i = 0; i < 9; ++i) Rd[i] = (best & 1) ? pd.R2[i] : pd.R1[i];
                for (int i = 0; i < 3; ++i) td[i] = (best & 2) ? -pd.t[i] : pd.t[i];
                pose_errors_deg<D>(Rd, td, Rgd, tgd, erd, etd);
                for (int i = 0; i < 9; ++i) {
                    const double gval = 0.5 * (erd.d[i] + etd.d[i]);
                    grad[bm * 9 + i] = (gval == gval && fabs(gval) < 1e30) ? (float)gval : 0.f;
                }
            }
        }
    }
    if (!mk) return;
    __syncthreads();
    const int best = best_s;
    for (int n = threadIdx.x; n < N; n += kPoseThreads) mk[n] = (mk[n] >> best) & 1;  // own writes only
}

}  // namespace drb

extern "C" int drb_recover_pose(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                                const float* t_gt, int B, int M, int N, float dist, float* R, float* t, uint8_t* mask,
                                int32_t* ngood, float* err, void* stream) {
    if (!E || !matches || !R || !t || !ngood) return DRB_ERR_NULL_POINTER;
    if (err && (!R_gt || !t_gt)) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || M <= 0 || N <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    drb::recover_pose_kernel<false><<<dim3(M, B), drb::kPoseThreads, 0, (cudaStream_t)stream>>>(
        E, matches, npts, R_gt, t_gt, M, N, dist, R, t, mask, ngood, err, nullptr);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_pose_loss(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                             const float* t_gt, int B, int M, int N, float dist, float* err, float* grad,
                             void* stream) {
    if (!E || !matches || !R_gt || !t_gt || !err) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || M <= 0 || N <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    drb::recover_pose_kernel<true><<<dim3(M, B), drb::kPoseThreads, 0, (cudaStream_t)stream>>>(
        E, matches, npts, R_gt, t_gt, M, N, dist, nullptr, nullptr, nullptr, nullptr, err, grad);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
