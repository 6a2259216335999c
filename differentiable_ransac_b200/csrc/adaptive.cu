// The test-mode driver's chunked loop with adaptive termination, without host syncs.
//
// The reference scores `ransac_batch_size` hypotheses per trip of a Python `while`, keeps the best so far,
// and after every improvement shrinks the trip budget from the winner's inlier count (ransac.py:55-144,
// :202-215): one device->host sync per trip.  Here all max_iterations hypotheses of all pairs go through
// sample -> solve -> score in ONE pass (they are independent); what the loop adds is bookkeeping, replayed on
// the device from the per-model scores:
//   chunk_argmax  best (score, id) of every chunk of `span` consecutive model ids   (torch.argmax, :114)
//   chunk_ninl    inlier count of every chunk winner                                 (best_mask.sum(), :120)
//   adaptive_scan one thread per pair walks its chunks in order, applies `>` (:116) and
//                 adaptive_iteration_number (:202-215) in double, stops where the reference stops
// Result: the winner the reference's loop would return and its `iterations`, for B pairs at once.  Chunks past
// the stopping point were computed for nothing -- at ~10 ns per hypothesis that is cheaper than one sync.
#include <cuda_runtime.h>

#include <cmath>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "sampson.cuh"

namespace drb {

__global__ void chunk_argmax_kernel(const float* __restrict__ scores, const int32_t* __restrict__ count,
                                    const int32_t* __restrict__ ids, int M, int span, int C,
                                    unsigned long long* __restrict__ chunk_best) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = count ? min(count[b], M) : M;
    if (j >= cnt) return;
    const int id = ids ? ids[(size_t)b * M + j] : j;
    const int c = id / span;
    if (c >= C) return;
    const unsigned long long key = pack_best(scores[(size_t)b * M + j], id);
    if (key) atomicMax(chunk_best + (size_t)b * C + c, key);
}

__global__ void __launch_bounds__(128)
chunk_ninl_kernel(const float* __restrict__ matches, const float* __restrict__ models_dense,
                  const unsigned long long* __restrict__ chunk_best, const float* __restrict__ thr, int Md, int N,
                  int C, int32_t* __restrict__ chunk_ninl) {
    __shared__ int warp_cnt[4];
    const int c = blockIdx.x, b = blockIdx.y;
    const int id = packed_id(chunk_best[(size_t)b * C + c]);
    float m[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i)  // nothing scored in this chunk: the identity, like best_finalize
        m[i] = (id >= 0 && id < Md) ? models_dense[((size_t)b * Md + id) * 9 + i] : ((i % 4 == 0) ? 1.f : 0.f);
    const float t = 1.5f * thr[b];
    const float thr2 = t * t;
    int n_in = 0;
    for (int n = threadIdx.x; n < N; n += 128) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)b * N + n);
        const Sampson s = sampson(m, p.x, p.y, p.z, p.w);
        n_in += (__fdiv_rn(s.r * s.r, s.j) < thr2) ? 1 : 0;  // the same test as the winner mask (msac_score.py:47)
    }
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = n_in;
    __syncthreads();
    if (threadIdx.x == 0) chunk_ninl[(size_t)b * C + c] = warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
}

// ransac.py:202-215, in double like the Python floats it is written in.
__device__ double adaptive_iteration_number(int ninl, int N, double confidence, int sample_size, int max_iterations,
                                            double eps) {
    const double ratio = (double)ninl / (double)N;
    const double p = pow(ratio, (double)sample_size);
    if (1.0 - p >= 1.0 - eps) return (double)max_iterations;
    const double v = log10(1.0 - confidence) / log10(1.0 - p + eps);
    return v > 0.0 ? v : 0.0;
}

__global__ void adaptive_scan_kernel(const unsigned long long* __restrict__ chunk_best,
                                     const int32_t* __restrict__ chunk_ninl, int B, int C, int N, int rbs,
                                     int max_iterations, int sample_size, double confidence, double eps,
                                     unsigned long long* __restrict__ best_packed, int32_t* __restrict__ iterations) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    unsigned long long best = 0ull;
    float best_score = 0.f;
    double max_iters = (double)max_iterations;
    int it = 0;
    for (int c = 0; c < C && (double)it < max_iters; ++c) {
        const unsigned long long key = chunk_best[(size_t)b * C + c];
        const float sc = packed_score(key);
        if (sc > best_score || it == 0) {  // ransac.py:116
            best_score = sc;
            best = key;
            const double a = adaptive_iteration_number(chunk_ninl[(size_t)b * C + c], N, confidence, sample_size,
                                                       max_iterations, eps);
            max_iters = a < (double)max_iterations ? a : (double)max_iterations;
        }
        it += rbs;
    }
    best_packed[b] = best;
    iterations[b] = it;
}

}  // namespace drb

extern "C" int drb_adaptive_select(const float* matches, const float* models_dense, const float* scores,
                                   const int32_t* count, const int32_t* ids, const float* thr, int B, int M, int Md,
                                   int N, int span, int rbs, int max_iterations, int sample_size, double confidence,
                                   double eps, unsigned long long* chunk_best, int32_t* chunk_ninl,
                                   unsigned long long* best_packed, int32_t* iterations, void* stream) {
    if (!matches || !models_dense || !scores || !thr || !chunk_best || !chunk_ninl || !best_packed || !iterations)
        return DRB_ERR_NULL_POINTER;
    if (B <= 0 || M <= 0 || Md <= 0 || N <= 0 || span <= 0 || rbs <= 0 || max_iterations <= 0 || sample_size <= 0 ||
        B > 65535)
        return DRB_ERR_BAD_SHAPE;
    const int C = (max_iterations + rbs - 1) / rbs;
    if (C > 65535 || (long long)C * span < (long long)Md) return DRB_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(chunk_best, 0, sizeof(unsigned long long) * (size_t)B * C, st);
    drb::chunk_argmax_kernel<<<dim3((M + 255) / 256, B), 256, 0, st>>>(scores, count, ids, M, span, C, chunk_best);
    drb::chunk_ninl_kernel<<<dim3(C, B), 128, 0, st>>>(matches, models_dense, chunk_best, thr, Md, N, C, chunk_ninl);
    drb::adaptive_scan_kernel<<<(B + 127) / 128, 128, 0, st>>>(chunk_best, chunk_ninl, B, C, N, rbs, max_iterations,
                                                               sample_size, confidence, eps, best_packed, iterations);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
