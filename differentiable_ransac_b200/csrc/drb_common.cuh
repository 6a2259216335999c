// Common definitions for the differentiable-RANSAC kernels (sm_100a).
//
// All numerical routines in *_math.cuh are written as DRB_HD templates so that
// (a) the CUDA kernels instantiate them in fp32 on the device and (b) the CPU
// test-suite can compile the very same arithmetic for the host
// (tests/hostcheck) and compare it with the oracle without a GPU.  The host
// instantiation is test infrastructure only; the Python package never loads it.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DRB_HD __host__ __device__ __forceinline__
#define DRB_D __device__ __forceinline__
#else
#define DRB_HD inline
#define DRB_D inline
#endif

#ifndef DRB_UNROLL
#if defined(__CUDACC__)
#define DRB_UNROLL _Pragma("unroll")
#else
#define DRB_UNROLL
#endif
#endif

namespace drb {

template <class T> DRB_HD T t_abs(T x) { return x < T(0) ? -x : x; }
template <class T> DRB_HD T t_max(T a, T b) { return a > b ? a : b; }
template <class T> DRB_HD T t_min(T a, T b) { return a < b ? a : b; }
// float / double: one instruction each on the device (|x| is an operand modifier, max is FMNMX); they differ from
// the generic forms only when an argument is NaN (fmax returns the other argument)
DRB_HD float t_abs(float x) { return fabsf(x); }
DRB_HD double t_abs(double x) { return fabs(x); }
DRB_HD float t_max(float a, float b) { return fmaxf(a, b); }
DRB_HD double t_max(double a, double b) { return fmax(a, b); }
// 1 / x where one ulp does not matter (scale factors, pivots of an elimination whose result is polished, Newton
// steps): MUFU.RCP on the device instead of the ~10-instruction IEEE division; exact elsewhere
template <class T> DRB_HD T t_rcp(T x) { return T(1) / x; }
DRB_HD float t_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
DRB_HD float t_sqrt(float x) { return sqrtf(x); }
DRB_HD double t_sqrt(double x) { return sqrt(x); }
DRB_HD float t_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
DRB_HD double t_rsqrt(double x) { return 1.0 / sqrt(x); }

// 3x3 helpers, row-major m[9]
template <class T> DRB_HD T det3(const T* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
template <class T> DRB_HD void cofactor3(const T* m, T* c) {  // c_ij = d det / d m_ij
    c[0] = m[4] * m[8] - m[5] * m[7];
    c[1] = m[5] * m[6] - m[3] * m[8];
    c[2] = m[3] * m[7] - m[4] * m[6];
    c[3] = m[2] * m[7] - m[1] * m[8];
    c[4] = m[0] * m[8] - m[2] * m[6];
    c[5] = m[1] * m[6] - m[0] * m[7];
    c[6] = m[1] * m[5] - m[2] * m[4];
    c[7] = m[2] * m[3] - m[0] * m[5];
    c[8] = m[0] * m[4] - m[1] * m[3];
}
// C = A * B
template <class T> DRB_HD void mul33(const T* a, const T* b, T* c) {
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    }
}
// C = A * B^T
template <class T> DRB_HD void mul33_nt(const T* a, const T* b, T* c) {
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
    }
}
// C = A^T * B
template <class T> DRB_HD void mul33_tn(const T* a, const T* b, T* c) {
    DRB_UNROLL
    for (int i = 0; i < 3; ++i) {
        DRB_UNROLL
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
    }
}

}  // namespace drb
