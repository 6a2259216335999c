// The hypothesize-and-score path in double precision (`-pr 2`: utils.py:42, model_cl.py:164-170).
//
// In the reference the precision flag sets the dtype of the sampler's soft one-hot; the minimal samples
// (ransac.py:64-65: matches * one-hot), the five-point solver (nister.py:121-122 takes the dtype of its input) and
// MSAC (msac_score.py:12-55) then run in float64 by type promotion.  The fp32 kernels of this library are what
// BASELINE.json times; this file is the same chain for callers who ask for float64: every hypothesis solved with
// the SAME templated math as the fp32 kernels and the host build (e5_math.cuh / poly_roots.cuh instantiated in
// double, serial per thread), every model scored in double, arg-max and winner mask in double.  Built for
// results, not for the roofline: one thread per hypothesis with its 10 x 20 system in local memory, one thread per
// model over all correspondences (B200 runs FP64 at half the FP32 rate).  Against the reference run in float64 the
// genuine models agree to ~1e-9 and the scores to ~1e-12 (tests/test_gpu_fp64_path.py).
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "e5_math.cuh"
#include "refit_math.cuh"

namespace drb {

struct F64Sink {
    double* dst;
    __device__ __forceinline__ void operator()(int slot, int i, double v) { dst[slot * 9 + i] = v; }
};

constexpr int kF64SolveThreads = 64;

__global__ void __launch_bounds__(kF64SolveThreads)
solve_e5_f64_kernel(const double* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                    double* __restrict__ models, int32_t* __restrict__ nsol) {
    const long long row = (long long)blockIdx.x * kF64SolveThreads + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    double p[5][4];
    for (int j = 0; j < 5; ++j) {
        const double* src = idx ? matches + ((size_t)b * N + idx[row * 5 + j]) * 4 : matches + (row * 5 + j) * 4;
        for (int c = 0; c < 4; ++c) p[j][c] = src[c];
    }
    LocalMat<double> M;
    double* dst = models + row * 90;
    F64Sink sink{dst};
    const int n = e5_solve<double, LocalMat<double>, double, F64Sink>(p, M, sink, 2);
    for (int s = n; s < 10; ++s)                       // identity in the unused slots (nister.py:400-401)
        for (int i = 0; i < 9; ++i) dst[s * 9 + i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    nsol[row] = n;
}

struct SampsonD {
    double r, j;
};
__device__ __forceinline__ SampsonD sampson_d(const double* m, double x1, double y1, double x2, double y2) {
    const double e0 = m[0] * x1 + m[1] * y1 + m[2], e1 = m[3] * x1 + m[4] * y1 + m[5], e2 = m[6] * x1 + m[7] * y1 + m[8];
    const double f0 = m[0] * x2 + m[3] * y2 + m[6], f1 = m[1] * x2 + m[4] * y2 + m[7];
    SampsonD s;
    s.r = x2 * e0 + y2 * e1 + e2;
    s.j = e0 * e0 + e1 * e1 + f0 * f0 + f1 * f1;
    return s;
}

constexpr int kF64ScoreThreads = 128;

// models[B, K * slots, 9]; slot s of hypothesis k is scored when s < nsol[b, k] (nsol nullable: every model);
// scores[B, K * slots], -1 for a slot that holds no model.
__global__ void __launch_bounds__(kF64ScoreThreads)
score_msac_f64_kernel(const double* __restrict__ matches, const double* __restrict__ models, const int32_t* __restrict__ nsol,
                      const double* __restrict__ thr, int K, int slots, int N, double* __restrict__ scores) {
    const int b = blockIdx.y;
    const int M = K * slots;
    const int m = blockIdx.x * kF64ScoreThreads + threadIdx.x;
    if (m >= M) return;
    const int k = m / slots, s = m - k * slots;
    if (nsol && s >= nsol[(size_t)b * K + k]) {
        scores[(size_t)b * M + m] = -1.0;
        return;
    }
    double mm[9];
    for (int i = 0; i < 9; ++i) mm[i] = models[((size_t)b * M + m) * 9 + i];
    const double th = 1.5 * thr[b];
    const double inv = 1.0 / (th * th);
    const double* pts = matches + (size_t)b * N * 4;
    double acc = 0.0;
    for (int n = 0; n < N; ++n) {
        const double2 a = *reinterpret_cast<const double2*>(pts + 4 * n), c = *reinterpret_cast<const double2*>(pts + 4 * n + 2);
        const SampsonD sp = sampson_d(mm, a.x, a.y, c.x, c.y);
        const double v = 1.0 - (sp.r * sp.r / sp.j) * inv;
        acc += v > 0.0 ? v : 0.0;                    // NaN compares false: contributes 0, like a clamped outlier
    }
    scores[(size_t)b * M + m] = acc;
}

// One CTA per pair: arg-max of the scores (first maximum, like torch.argmax at ransac.py:114), the winner's model,
// score, inlier mask (d2 < (1.5 thr)^2, msac_score.py:44) and inlier count.
__global__ void __launch_bounds__(256)
best_finalize_f64_kernel(const double* __restrict__ matches, const double* __restrict__ models, const double* __restrict__ scores,
                         const double* __restrict__ thr, int M, int N, int32_t* __restrict__ best_id,
                         double* __restrict__ best_score, double* __restrict__ best_model, uint8_t* __restrict__ mask,
                         int32_t* __restrict__ ninl) {
    __shared__ double sv[256];
    __shared__ int si[256];
    __shared__ int cnt[8];
    const int b = blockIdx.x, t = threadIdx.x;
    double bv = -1.0;
    int bi = -1;
    for (int m = t; m < M; m += 256) {
        const double v = scores[(size_t)b * M + m];
        if (v > bv) { bv = v; bi = m; }             // strict: the lowest index among equals within a thread
    }
    sv[t] = bv;
    si[t] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (t < o) {
            const double ov = sv[t + o];
            const int oi = si[t + o];
            if (oi >= 0 && (si[t] < 0 || ov > sv[t] || (ov == sv[t] && oi < si[t]))) { sv[t] = ov; si[t] = oi; }
        }
        __syncthreads();
    }
    const int id = si[0];
    double mm[9];
    for (int i = 0; i < 9; ++i) mm[i] = id >= 0 ? models[((size_t)b * M + id) * 9 + i] : ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0);
    if (t == 0) {
        best_id[b] = id;
        best_score[b] = id >= 0 ? sv[0] : 0.0;
        for (int i = 0; i < 9; ++i) best_model[b * 9 + i] = mm[i];
    }
    const double th = 1.5 * thr[b], th2 = th * th;
    int c = 0;
    for (int n = t; n < N; n += 256) {
        const double* p = matches + ((size_t)b * N + n) * 4;
        const SampsonD sp = sampson_d(mm, p[0], p[1], p[2], p[3]);
        const bool in = id >= 0 && (sp.r * sp.r / sp.j) < th2;
        mask[(size_t)b * N + n] = in ? 1 : 0;
        c += in ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((t & 31) == 0) cnt[t >> 5] = c;
    __syncthreads();
    if (t == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += cnt[w];
        ninl[b] = tot;
    }
}

}  // namespace drb

using namespace drb;

extern "C" int drb_solve_e5_f64(const double* matches, const int32_t* idx, int B, int K, int N, double* models,
                                int32_t* nsol, void* stream) {
    if (!matches || !models || !nsol) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    const long long rows = (long long)B * K;
    solve_e5_f64_kernel<<<(unsigned)((rows + kF64SolveThreads - 1) / kF64SolveThreads), kF64SolveThreads, 0,
                          (cudaStream_t)stream>>>(matches, idx, B, K, N, models, nsol);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_score_msac_f64(const double* matches, const double* models, const int32_t* nsol, const double* thr, int B,
                                  int K, int slots, int N, double* scores, void* stream) {
    if (!matches || !models || !thr || !scores) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || B > 65535 || K <= 0 || slots <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    const int M = K * slots;
    score_msac_f64_kernel<<<dim3((M + kF64ScoreThreads - 1) / kF64ScoreThreads, B), kF64ScoreThreads, 0, (cudaStream_t)stream>>>(
        matches, models, nsol, thr, K, slots, N, scores);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_best_finalize_f64(const double* matches, const double* models, const double* scores, const double* thr,
                                     int B, int M, int N, int32_t* best_id, double* best_score, double* best_model,
                                     uint8_t* mask, int32_t* ninl, void* stream) {
    if (!matches || !models || !scores || !thr || !best_id || !best_score || !best_model || !mask || !ninl)
        return DRB_ERR_NULL_POINTER;
    if (B <= 0 || M <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    best_finalize_f64_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(matches, models, scores, thr, M, N, best_id, best_score,
                                                                  best_model, mask, ninl);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
