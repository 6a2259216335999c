// Pose recovery on the device, two launches:
//   pose_vote_kernel    one THREAD per (correspondence, candidate pose): DLT triangulation in double, the
//                       cheirality test, warp-ballot -> per-pose counts (atomics) and a 4-bit code per point
//   pose_finish_kernel  one CTA per (model, pair): arg-max of the counts, the winning pose, its mask, and --
//                       given a ground-truth pose -- the angular errors (and, for PoseLoss, their gradient
//                       with respect to E by forward-mode duals)
// Every CTA re-derives the decomposition of its model (a 3 x 3 Jacobi: cheaper than a round trip through
// global memory and a third launch).  See pose_math.cuh for the reference lines this replaces.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "pose_math.cuh"

namespace drb {

constexpr int kPoseThreads = 128;

// HORN selects the decomposition: false = SVD-equivalent (cv_utils.decompose_E), true = Horn's closed form
// (cv_utils.new_decompose_E, what PoseLoss differentiates).
template <bool HORN>
__device__ __forceinline__ bool decompose_model(const float* __restrict__ E, PoseCandidates<double>& pc) {
    double e[9];
    bool fin = true;
    for (int i = 0; i < 9; ++i) {
        e[i] = (double)E[i];
        fin = fin && (e[i] == e[i]) && (fabs(e[i]) < 1e30);
    }
    if (!fin) return false;
    return HORN ? decompose_essential_horn<double>(e, pc) : decompose_essential<double>(e, pc);
}

template <bool HORN>
__global__ void __launch_bounds__(kPoseThreads)
pose_vote_kernel(const float* __restrict__ E, const float* __restrict__ matches, const int32_t* __restrict__ npts,
                 int M, int N, float dist, uint8_t* __restrict__ codes, int32_t* __restrict__ counts) {
    __shared__ PoseCandidates<double> pc;
    __shared__ int ok_s;
    const int m = blockIdx.y, b = blockIdx.z;
    const size_t bm = (size_t)b * M + m;
    if (threadIdx.x == 0) ok_s = decompose_model<HORN>(E + bm * 9, pc) ? 1 : 0;
    __syncthreads();
    if (!ok_s) return;  // counts stay zero; the finish kernel writes the failure outputs
    const int g = blockIdx.x * kPoseThreads + threadIdx.x;
    const int n = g >> 2, c = g & 3;
    const int n_used = npts ? min(npts[b], N) : N;
    bool in_front = false;
    if (n < n_used) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)b * N + n);
        in_front = cheirality_one<double>(pc, c, (double)p.x, (double)p.y, (double)p.z, (double)p.w, (double)dist);
    }
    const unsigned ball = __ballot_sync(0xffffffffu, in_front);
    const int lane = threadIdx.x & 31;
    if (lane < 4) {  // pose `lane` sits in lanes lane, lane + 4, ...
        const int k = __popc(ball & (0x11111111u << lane));
        if (k) atomicAdd(counts + bm * 4 + lane, k);
    }
    if (codes && c == 0 && n < N) codes[bm * N + n] = (uint8_t)((ball >> lane) & 0xFu);
}

template <bool HORN>
__global__ void __launch_bounds__(kPoseThreads)
pose_finish_kernel(const float* __restrict__ E, const float* __restrict__ R_gt, const float* __restrict__ t_gt,
                   const int32_t* __restrict__ counts, int M, int N, float* __restrict__ R, float* __restrict__ t,
                   uint8_t* __restrict__ mask, int32_t* __restrict__ ngood, float* __restrict__ err,
                   float* __restrict__ grad) {
    __shared__ int best_s;
    const int m = blockIdx.x, b = blockIdx.y;
    const size_t bm = (size_t)b * M + m;
    if (threadIdx.x == 0) {
        PoseCandidates<double> pc;
        int best = -1;
        if (decompose_model<HORN>(E + bm * 9, pc)) {
            best = 0;
            for (int c = 1; c < 4; ++c)
                if (counts[bm * 4 + c] > counts[bm * 4 + best]) best = c;  // first maximum (cv_utils.py:71)
        }
        best_s = best;
        if (best < 0) {  // not decomposable: identity pose, nothing in front of anything
            if (R)
                for (int i = 0; i < 9; ++i) R[bm * 9 + i] = (i % 4 == 0) ? 1.f : 0.f;
            if (t)
                for (int i = 0; i < 3; ++i) t[bm * 3 + i] = 0.f;
            if (ngood) ngood[bm] = 0;
            if (err) { err[bm * 2] = 180.f; err[bm * 2 + 1] = 90.f; }  // eval_essential_matrix's failure values
            if (grad)
                for (int i = 0; i < 9; ++i) grad[bm * 9 + i] = 0.f;
        } else {
            double Rb[9], tb[3];
            for (int i = 0; i < 9; ++i) Rb[i] = (best & 1) ? pc.R2[i] : pc.R1[i];
            for (int i = 0; i < 3; ++i) tb[i] = (best & 2) ? -pc.t[i] : pc.t[i];
            if (R)
                for (int i = 0; i < 9; ++i) R[bm * 9 + i] = (float)Rb[i];
            if (t)
                for (int i = 0; i < 3; ++i) t[bm * 3 + i] = (float)tb[i];
            if (ngood) ngood[bm] = counts[bm * 4 + best];
            if (err && R_gt && t_gt) {
                double Rg[9], tg[3], er, et;
                for (int i = 0; i < 9; ++i) Rg[i] = (double)R_gt[b * 9 + i];
                for (int i = 0; i < 3; ++i) tg[i] = (double)t_gt[b * 3 + i];
                pose_errors_deg<double>(Rb, tb, Rg, tg, er, et);
                err[bm * 2] = (float)er;
                err[bm * 2 + 1] = (float)et;
                if (HORN && grad) {
                    // the closed form once more on duals: d((err_R + err_t) / 2)/dE, one PoseLoss term's backward
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
    }
    if (!mask) return;
    __syncthreads();
    const int best = best_s;
    uint8_t* mk = mask + bm * N;  // holds the 4-bit codes of the vote; becomes the winner's mask in place
    for (int n = threadIdx.x; n < N; n += kPoseThreads) mk[n] = best < 0 ? 0 : ((mk[n] >> best) & 1);
}

template <bool HORN>
static int launch_pose(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                       const float* t_gt, int B, int M, int N, float dist, int32_t* counts, float* R, float* t,
                       uint8_t* mask, int32_t* ngood, float* err, float* grad, cudaStream_t st) {
    if (B <= 0 || M <= 0 || N <= 0 || B > 65535 || M > 65535) return DRB_ERR_BAD_SHAPE;
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * M * 4, st);
    const int gx = (4 * N + kPoseThreads - 1) / kPoseThreads;
    pose_vote_kernel<HORN><<<dim3(gx, M, B), kPoseThreads, 0, st>>>(E, matches, npts, M, N, dist, mask, counts);
    pose_finish_kernel<HORN><<<dim3(M, B), kPoseThreads, 0, st>>>(E, R_gt, t_gt, counts, M, N, R, t, mask, ngood, err,
                                                                  grad);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

}  // namespace drb

extern "C" int drb_recover_pose(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                                const float* t_gt, int B, int M, int N, float dist, int32_t* counts, float* R,
                                float* t, uint8_t* mask, int32_t* ngood, float* err, void* stream) {
    if (!E || !matches || !counts || !R || !t || !ngood) return DRB_ERR_NULL_POINTER;
    if (err && (!R_gt || !t_gt)) return DRB_ERR_NULL_POINTER;
    return drb::launch_pose<false>(E, matches, npts, R_gt, t_gt, B, M, N, dist, counts, R, t, mask, ngood, err, nullptr,
                                   (cudaStream_t)stream);
}

extern "C" int drb_pose_loss(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                             const float* t_gt, int B, int M, int N, float dist, int32_t* counts, float* err,
                             float* grad, void* stream) {
    if (!E || !matches || !R_gt || !t_gt || !counts || !err) return DRB_ERR_NULL_POINTER;
    return drb::launch_pose<true>(E, matches, npts, R_gt, t_gt, B, M, N, dist, counts, nullptr, nullptr, nullptr,
                                  nullptr, err, grad, (cudaStream_t)stream);
}
