// Five-point essential-matrix kernels: one hypothesis per thread, the 10 x 20
// constraint matrix of every thread resident in shared memory (column-interleaved,
// bank-conflict free), everything else in registers.  See e5_math.cuh for the math
// and the reference lines it replaces.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "e5_math.cuh"
#include "e5_backward.cuh"

namespace drb {

constexpr int kE5Threads = 64;
constexpr int kE5SmemBytes = kE5Threads * 200 * sizeof(float);

struct SmemMat {
    float* base;  // smem + tid ; element (r, c) at base[(r * 20 + c) * kE5Threads]
    __device__ __forceinline__ float& operator()(int r, int c) { return base[(r * 20 + c) * kE5Threads]; }
};

struct SmemSink {
    float* base;  // same column as SmemMat: element e (= slot * 9 + i, e < 90 <= 200) at base[e * kE5Threads]
    __device__ __forceinline__ void operator()(int slot, int i, float v) { base[(slot * 9 + i) * kE5Threads] = v; }
    __device__ __forceinline__ float get(int e) const { return base[e * kE5Threads]; }
};

__device__ __forceinline__ void load_minimal5(const float* __restrict__ matches, const int32_t* __restrict__ idx,
                                              long long row, int b, int N, float (*p)[4]) {
    DRB_UNROLL
    for (int j = 0; j < 5; ++j) {
        const float4* src;
        if (idx != nullptr) {
            src = reinterpret_cast<const float4*>(matches) + (size_t)b * N + idx[row * 5 + j];
        } else {
            src = reinterpret_cast<const float4*>(matches) + row * 5 + j;
        }
        const float4 v = __ldg(src);
        p[j][0] = v.x; p[j][1] = v.y; p[j][2] = v.z; p[j][3] = v.w;
    }
}

__global__ void __launch_bounds__(kE5Threads)
solve_e5_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                float* __restrict__ models, int32_t* __restrict__ nsol, float* __restrict__ cmodels,
                int32_t* __restrict__ cids, int32_t* __restrict__ ccount) {
    extern __shared__ float smem[];
    const long long row = (long long)blockIdx.x * kE5Threads + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    const int k = (int)(row % K);
    float p[5][4];
    load_minimal5(matches, idx, row, b, N, p);
    SmemMat M{smem + threadIdx.x};
    // The solutions are staged in this thread's own scratch column (dead once the z-polynomials have
    // been read out), so the dense and the compact copies below read shared memory, not the global
    // memory that was just written.
    SmemSink sink{smem + threadIdx.x};
    const int n = e5_solve<float, SmemMat, float, SmemSink>(p, M, sink, 2);
    nsol[row] = n;
    // reserve the compact-list slots early so the atomic's round trip overlaps the dense write
    int pos = 0;
    if (cmodels != nullptr && n > 0) pos = atomicAdd(ccount + b, n);
    float2* dense = reinterpret_cast<float2*>(models + (size_t)row * 90);   // 360 B per row: 8-byte aligned
    DRB_UNROLL
    for (int e = 0; e < 45; ++e) {
        const int e0 = 2 * e, e1 = 2 * e + 1;
        const float v0 = (e0 / 9 < n) ? sink.get(e0) : (((e0 % 9) % 4 == 0) ? 1.f : 0.f);
        const float v1 = (e1 / 9 < n) ? sink.get(e1) : (((e1 % 9) % 4 == 0) ? 1.f : 0.f);
        dense[e] = make_float2(v0, v1);
    }
    if (cmodels != nullptr && n > 0) {
        float* dst = cmodels + ((size_t)b * K * 10 + pos) * 9;
        for (int s = 0; s < n; ++s) {
            DRB_UNROLL
            for (int i = 0; i < 9; ++i) dst[s * 9 + i] = sink.get(s * 9 + i);
            cids[(size_t)b * K * 10 + pos + s] = k * 10 + s;
        }
    }
}

__global__ void __launch_bounds__(128)
solve_e5_backward_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                         const float* __restrict__ models, const int32_t* __restrict__ sel,
                         const float* __restrict__ g_model, float* __restrict__ g_pts) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    float gp[5][4];
    const int s = sel[row];
    bool ok = s >= 0;
    if (ok) {
        float p[5][4], E[9], g[9];
        load_minimal5(matches, idx, row, b, N, p);
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) {
            E[i] = models[(size_t)row * 90 + s * 9 + i];
            g[i] = g_model[(size_t)row * 9 + i];
        }
        ok = e5_backward<float>(p, E, g, gp);
    }
    DRB_UNROLL
    for (int j = 0; j < 5; ++j) {
        float4 o = ok ? make_float4(gp[j][0], gp[j][1], gp[j][2], gp[j][3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(g_pts)[row * 5 + j] = o;
    }
}

__global__ void select_closest_kernel(const float* __restrict__ models, const int32_t* __restrict__ nsol,
                                      const float* __restrict__ gt, int B, int K, int slots, int sign_invariant,
                                      int32_t* __restrict__ sel, float* __restrict__ chosen) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    float g[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) g[i] = gt[b * 9 + i];
    const int n = nsol[row];
    int best = -1;
    float bd = INFINITY;
    for (int s = 0; s < n && s < slots; ++s) {
        float dp = 0.f, dn = 0.f;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) {
            const float m = models[((size_t)row * slots + s) * 9 + i];
            dp += (m - g[i]) * (m - g[i]);
            dn += (m + g[i]) * (m + g[i]);
        }
        const float d = sign_invariant ? fminf(dp, dn) : dp;
        if (d < bd) { bd = d; best = s; }
    }
    sel[row] = best;
    DRB_UNROLL
    for (int i = 0; i < 9; ++i)
        chosen[(size_t)row * 9 + i] = best >= 0 ? models[((size_t)row * slots + best) * 9 + i]
                                                : ((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
}

}  // namespace drb

using namespace drb;

extern "C" int drb_solve_e5(const float* matches, const int32_t* idx, int B, int K, int N, float* models,
                            int32_t* nsol, float* cmodels, int32_t* cids, int32_t* ccount, void* stream) {
    if (!matches || !models || !nsol) return DRB_ERR_NULL_POINTER;
    if (cmodels && (!cids || !ccount)) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(solve_e5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kE5SmemBytes) !=
            cudaSuccess)
            return DRB_ERR_CUDA;
        configured = true;
    }
    const long long rows = (long long)B * K;
    solve_e5_kernel<<<(unsigned)((rows + kE5Threads - 1) / kE5Threads), kE5Threads, kE5SmemBytes,
                      (cudaStream_t)stream>>>(matches, idx, B, K, N, models, nsol, cmodels, cids, ccount);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_solve_e5_backward(const float* matches, const int32_t* idx, int B, int K, int N,
                                     const float* models, const int32_t* sel, const float* g_model, float* g_pts,
                                     void* stream) {
    if (!matches || !models || !sel || !g_model || !g_pts) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    const long long rows = (long long)B * K;
    solve_e5_backward_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        matches, idx, B, K, N, models, sel, g_model, g_pts);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_select_closest(const float* models, const int32_t* nsol, const float* gt, int B, int K, int slots,
                                  int sign_invariant, int32_t* sel, float* chosen, void* stream) {
    if (!models || !nsol || !gt || !sel || !chosen) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || slots <= 0) return DRB_ERR_BAD_SHAPE;
    const long long rows = (long long)B * K;
    select_closest_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        models, nsol, gt, B, K, slots, sign_invariant, sel, chosen);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
