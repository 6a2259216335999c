// Fundamental-matrix (8-point, 7-point) and rigid 3-point kernels, one hypothesis per
// thread, everything in registers.  Math and reference citations: f8_math.cuh, rigid_math.cuh.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "f8_math.cuh"
#include "rigid_math.cuh"

namespace drb {

template <int S>
__device__ __forceinline__ void load_minimal2d(const float* __restrict__ matches, const int32_t* __restrict__ idx,
                                               long long row, int b, int N, float (*p)[4]) {
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        const float4* src = idx ? reinterpret_cast<const float4*>(matches) + (size_t)b * N + idx[row * S + j]
                                : reinterpret_cast<const float4*>(matches) + row * S + j;
        const float4 v = __ldg(src);
        p[j][0] = v.x; p[j][1] = v.y; p[j][2] = v.z; p[j][3] = v.w;
    }
}

__device__ __forceinline__ void load_minimal3d(const float* __restrict__ points, const int32_t* __restrict__ idx,
                                               long long row, int b, int N, float (*p)[6]) {
    DRB_UNROLL
    for (int j = 0; j < 3; ++j) {
        const float2* src = idx ? reinterpret_cast<const float2*>(points) + ((size_t)b * N + idx[row * 3 + j]) * 3
                                : reinterpret_cast<const float2*>(points) + (row * 3 + j) * 3;
        const float2 a = __ldg(src), c = __ldg(src + 1), e = __ldg(src + 2);
        p[j][0] = a.x; p[j][1] = a.y; p[j][2] = c.x; p[j][3] = c.y; p[j][4] = e.x; p[j][5] = e.y;
    }
}

__global__ void __launch_bounds__(128)
solve_f8_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                float* __restrict__ models, uint8_t* __restrict__ valid) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    float p[8][4], F[9];
    load_minimal2d<8>(matches, idx, row, (int)(row / K), N, p);
    const bool ok = f8_solve<float>(p, F);
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) models[row * 9 + i] = ok ? F[i] : ((i % 4 == 0) ? 1.f : 0.f);
    if (valid) valid[row] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(128)
solve_f8_backward_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                         const float* __restrict__ models, const float* __restrict__ g_model,
                         float* __restrict__ g_pts) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    float p[8][4], g[9], gp[8][4];
    load_minimal2d<8>(matches, idx, row, (int)(row / K), N, p);
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) g[i] = g_model[row * 9 + i];
    float Ff[9];
    if (models != nullptr) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) Ff[i] = models[row * 9 + i];
    }
    const bool ok = f8_backward<float, double>(p, g, gp, models != nullptr ? Ff : nullptr);
    DRB_UNROLL
    for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(g_pts)[row * 8 + j] =
            ok ? make_float4(gp[j][0], gp[j][1], gp[j][2], gp[j][3]) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(128)
solve_f7_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx, int B, int K, int N,
                float* __restrict__ models, int32_t* __restrict__ nsol) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    float p[7][4], F[3][9];
    load_minimal2d<7>(matches, idx, row, (int)(row / K), N, p);
    const int n = f7_solve<float>(p, F);
    DRB_UNROLL
    for (int s = 0; s < 3; ++s) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) models[(row * 3 + s) * 9 + i] = F[s][i];
    }
    nsol[row] = n;
}

__global__ void __launch_bounds__(128)
solve_rigid3_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int B, int K, int N, int flag,
                    float* __restrict__ models, uint8_t* __restrict__ valid) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    float p[3][6], m[16];
    load_minimal3d(points, idx, row, (int)(row / K), N, p);
    const bool ok = rigid3_solve<float>(p, flag, m);
    DRB_UNROLL
    for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(models)[row * 4 + i] = make_float4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
    if (valid) valid[row] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(128)
solve_rigid3_backward_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int B, int K, int N,
                             int flag, const float* __restrict__ g_model, float* __restrict__ g_pts) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    float p[3][6], g[16], gp[3][6];
    load_minimal3d(points, idx, row, (int)(row / K), N, p);
    DRB_UNROLL
    for (int i = 0; i < 16; ++i) g[i] = g_model[row * 16 + i];
    const bool ok = rigid3_backward<float, double>(p, flag, g, gp);
    DRB_UNROLL
    for (int j = 0; j < 3; ++j) {
        DRB_UNROLL
        for (int c = 0; c < 6; ++c) g_pts[(row * 3 + j) * 6 + c] = ok ? gp[j][c] : 0.f;
    }
}

}  // namespace drb

using namespace drb;

#define DRB_ROWS_LAUNCH(kernel, ...)                                                                \
    do {                                                                                            \
        const long long rows = (long long)B * K;                                                    \
        kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(__VA_ARGS__);      \
        return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;                           \
    } while (0)

extern "C" int drb_solve_f8(const float* matches, const int32_t* idx, int B, int K, int N, float* models,
                            uint8_t* valid, void* stream) {
    if (!matches || !models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    DRB_ROWS_LAUNCH(solve_f8_kernel, matches, idx, B, K, N, models, valid);
}

extern "C" int drb_solve_f8_backward(const float* matches, const int32_t* idx, int B, int K, int N,
                                     const float* models, const float* g_model, float* g_pts, void* stream) {
    // `models` (nullable): the forward's output, used only to fix the sign of the recomputed null vector
    if (!matches || !g_model || !g_pts) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    DRB_ROWS_LAUNCH(solve_f8_backward_kernel, matches, idx, B, K, N, models, g_model, g_pts);
}

extern "C" int drb_solve_f7(const float* matches, const int32_t* idx, int B, int K, int N, float* models,
                            int32_t* nsol, void* stream) {
    if (!matches || !models || !nsol) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    DRB_ROWS_LAUNCH(solve_f7_kernel, matches, idx, B, K, N, models, nsol);
}

extern "C" int drb_solve_rigid3(const float* points, const int32_t* idx, int B, int K, int N, int flag, float* models,
                                uint8_t* valid, void* stream) {
    if (!points || !models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    DRB_ROWS_LAUNCH(solve_rigid3_kernel, points, idx, B, K, N, flag, models, valid);
}

extern "C" int drb_solve_rigid3_backward(const float* points, const int32_t* idx, int B, int K, int N, int flag,
                                         const float* models, const float* g_model, float* g_pts, void* stream) {
    (void)models;
    if (!points || !g_model || !g_pts) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || (idx && N <= 0)) return DRB_ERR_BAD_SHAPE;
    DRB_ROWS_LAUNCH(solve_rigid3_backward_kernel, points, idx, B, K, N, flag, g_model, g_pts);
}
