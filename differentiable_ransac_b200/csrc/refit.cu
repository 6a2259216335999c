// Non-minimal refit kernels: one CTA per pair accumulates the 9 x 9 moment matrix of the selected
// correspondences in double, one thread then runs the serial tail of refit_math.cuh.  This is the step
// after the hypothesize-and-score loop (ransac.py:148-195, :217-257); it runs once per pair and call, so it
// is sized for latency, not throughput.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "refit_math.cuh"

namespace drb {

constexpr int kRefitThreads = 256;
constexpr int kRefitWarps = kRefitThreads / 32;

// Sum NV doubles per thread over the CTA; every thread gets the totals back in v[].
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double (*red)[45]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DRB_UNROLL
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
        DRB_UNROLL
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        DRB_UNROLL
        for (int w = 0; w < kRefitWarps; ++w) s += red[w][threadIdx.x];
        red[0][threadIdx.x] = s;
    }
    __syncthreads();
    DRB_UNROLL
    for (int i = 0; i < NV; ++i) v[i] = red[0][i];
    __syncthreads();
}

template <bool FMAT>
__global__ void __launch_bounds__(kRefitThreads)
refit_kernel(const float* __restrict__ matches, const uint8_t* __restrict__ mask, const float* __restrict__ weights,
             int N, float* __restrict__ models, int32_t* __restrict__ nsol) {
    constexpr int S = FMAT ? 1 : 10;
    constexpr int kMin = FMAT ? 8 : 5;
    __shared__ double red[kRefitWarps][45];
    const int b = blockIdx.x;
    const float4* m = reinterpret_cast<const float4*>(matches) + (size_t)b * N;
    const uint8_t* mk = mask ? mask + (size_t)b * N : nullptr;
    const float* w = weights ? weights + (size_t)b * N : nullptr;

    HartleyNorm<double> h;
    h.m[0] = h.m[1] = h.m[2] = h.m[3] = 0.0;
    h.r1 = h.r2 = 1.0;
    double count;
    {
        double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int n = threadIdx.x; n < N; n += kRefitThreads) {
            if (mk && !mk[n]) continue;
            s[0] += 1.0;
            if (FMAT) {
                const float4 p = __ldg(m + n);
                s[1] += p.x; s[2] += p.y; s[3] += p.z; s[4] += p.w;
            }
        }
        block_sum<5>(s, red);
        count = s[0];
        if (FMAT && count > 0.0) {
            DRB_UNROLL
            for (int c = 0; c < 4; ++c) h.m[c] = s[1 + c] / count;
        }
    }
    if (FMAT) {
        // fundamental_matrix_estimator.py:177-217: mean distance to the mass point -> sqrt(2)
        double d[2] = {0.0, 0.0};
        for (int n = threadIdx.x; n < N; n += kRefitThreads) {
            if (mk && !mk[n]) continue;
            const float4 p = __ldg(m + n);
            const double a = p.x - h.m[0], bb = p.y - h.m[1], c = p.z - h.m[2], e = p.w - h.m[3];
            d[0] += sqrt(a * a + bb * bb);
            d[1] += sqrt(c * c + e * e);
        }
        block_sum<2>(d, red);
        if (count > 0.0) {
            h.r1 = 1.4142135623730951 / (d[0] / count);
            h.r2 = 1.4142135623730951 / (d[1] / count);
        }
    }

    double acc[45];
    DRB_UNROLL
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    for (int n = threadIdx.x; n < N; n += kRefitThreads) {
        if (mk && !mk[n]) continue;
        const float4 p = __ldg(m + n);
        double row[9];
        epipolar_row<double>((p.x - h.m[0]) * h.r1, (p.y - h.m[1]) * h.r1, (p.z - h.m[2]) * h.r2,
                             (p.w - h.m[3]) * h.r2, row);
        // the reference scales the ROWS by the weight (nister.py:88-101, fundamental...:243-244): w^2 here
        const double ww = w ? (double)w[n] * (double)w[n] : 1.0;
        int e = 0;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) {
            const double ri = ww * row[i];
            DRB_UNROLL
            for (int j = i; j < 9; ++j) acc[e++] += ri * row[j];
        }
    }
    block_sum<45>(acc, red);

    if (threadIdx.x != 0) return;
    float* out = models + (size_t)b * S * 9;
    int n_out = 0;
    if (count >= (double)kMin) {
        if (FMAT) {
            double F[9];
            if (f8_refit_from_moments<double>(acc, h, F)) {
                for (int i = 0; i < 9; ++i) out[i] = (float)F[i];
                n_out = 1;
            }
        } else {
            double E[10][9];
            n_out = e5_refit_from_moments<double>(acc, E);
            for (int s = 0; s < n_out; ++s)
                for (int i = 0; i < 9; ++i) out[s * 9 + i] = (float)E[s][i];
        }
    }
    for (int s = n_out; s < S; ++s)  // identity padding, like nister.py:400-401
        for (int i = 0; i < 9; ++i) out[s * 9 + i] = (i % 4 == 0) ? 1.f : 0.f;
    nsol[b] = n_out;
}

}  // namespace drb

static int launch_refit(bool fmat, const float* matches, const uint8_t* mask, const float* weights, int B, int N,
                        float* models, int32_t* nsol, void* stream) {
    if (!matches || !models || !nsol) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (fmat) {
        drb::refit_kernel<true><<<B, drb::kRefitThreads, 0, st>>>(matches, mask, weights, N, models, nsol);
    } else {
        drb::refit_kernel<false><<<B, drb::kRefitThreads, 0, st>>>(matches, mask, weights, N, models, nsol);
    }
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

extern "C" int drb_refit_e5(const float* matches, const uint8_t* mask, const float* weights, int B, int N,
                            float* models, int32_t* nsol, void* stream) {
    return launch_refit(false, matches, mask, weights, B, N, models, nsol, stream);
}

extern "C" int drb_refit_f8(const float* matches, const uint8_t* mask, const float* weights, int B, int N,
                            float* models, int32_t* nsol, void* stream) {
    return launch_refit(true, matches, mask, weights, B, N, models, nsol, stream);
}
