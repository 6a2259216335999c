// Library-level entry points of libdrb.so (version, status strings, device facts)
// and the sparse gather backward.  See include/drb.h.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"

extern "C" int drb_version(void) { return 100; }

extern "C" const char* drb_status_string(int status) {
    switch (status) {
        case DRB_OK: return "ok";
        case DRB_ERR_NULL_POINTER: return "null pointer";
        case DRB_ERR_BAD_SHAPE: return "bad shape";
        case DRB_ERR_UNSUPPORTED: return "unsupported configuration";
        case DRB_ERR_CUDA: return "CUDA error";
        default: return "unknown status";
    }
}

extern "C" int drb_last_error(void) { return (int)cudaGetLastError(); }

extern "C" int drb_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

namespace drb {
// g_pts[B,K,s,D] -> g_sel[B,K,s] = <matches[n], g_pts>, grad_matches[B,N,D] += g_pts.
// (ret == 1 exactly on the selected entries, ransac.py:64, so the product rule gives these two.)
__global__ void gather_backward_kernel(const float* __restrict__ matches, const int32_t* __restrict__ idx,
                                       const float* __restrict__ g_pts, long long total, int K, int N, int s, int D,
                                       float* __restrict__ g_sel, float* __restrict__ grad_matches) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, k, j)
    if (e >= total) return;
    const int b = (int)(e / ((long long)K * s));
    const int n = idx[e];
    const float* m = matches + ((size_t)b * N + n) * D;
    const float* g = g_pts + (size_t)e * D;
    float acc = 0.f;
    for (int c = 0; c < D; ++c) {
        const float gv = g[c];
        acc += m[c] * gv;
        if (grad_matches) atomicAdd(grad_matches + ((size_t)b * N + n) * D + c, gv);
    }
    g_sel[e] = acc;
}
}  // namespace drb

extern "C" int drb_gather_backward(const float* matches, const int32_t* idx, const float* g_pts, int B, int K, int N,
                                   int s, int D, float* g_sel, float* grad_matches, void* stream) {
    if (!matches || !idx || !g_pts || !g_sel) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || s <= 0 || D <= 0) return DRB_ERR_BAD_SHAPE;
    const long long total = (long long)B * K * s;
    drb::gather_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        matches, idx, g_pts, total, K, N, s, D, g_sel, grad_matches);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

// Host -> device copy of a caller's buffer on a stream: cudaMemcpyAsync behind the C ABI, so that the pipelined
// service enqueues a batch's inputs with one foreign call per tensor instead of a framework dispatch each (the host
// side of a 0.14 ms step has ~100 us to spare in all).  Asynchronous when `src` is pinned.
extern "C" int drb_copy_h2d_async(void* dst_device, const void* src_host, size_t bytes, void* stream) {
    if (!dst_device || !src_host) return DRB_ERR_NULL_POINTER;
    if (bytes == 0) return DRB_OK;
    return cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream) == cudaSuccess
               ? DRB_OK
               : DRB_ERR_CUDA;
}

