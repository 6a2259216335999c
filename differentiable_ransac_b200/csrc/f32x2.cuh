// Packed two-lane fp32 arithmetic (PTX f32x2, SASS FFMA2 / FMUL2 / FADD2 -- new with sm_100):
// one issue slot performs two IEEE fp32 FMAs, which is what lifts the issue-bound residual
// kernels.  A pk2 holds (lo, hi) = the same quantity for two different correspondences.
#pragma once

#include <cuda_runtime.h>

namespace drb {

typedef unsigned long long pk2;

__device__ __forceinline__ pk2 pk2_make(float lo, float hi) {
    pk2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ pk2 pk2_splat(float v) { return pk2_make(v, v); }
__device__ __forceinline__ void pk2_split(pk2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pk2 pk2_fma(pk2 a, pk2 b, pk2 c) {
    pk2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pk2 pk2_mul(pk2 a, pk2 b) {
    pk2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// `volatile` twins: the compiler keeps volatile asm statements in program order, which lets a
// hand-written sequence put instructions that share a source operand next to each other (the
// register-file operand-reuse cache only helps adjacent instructions; score.cu).
__device__ __forceinline__ pk2 pk2_fma_v(pk2 a, pk2 b, pk2 c) {
    pk2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pk2 pk2_mul_v(pk2 a, pk2 b) {
    pk2 d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ pk2 pk2_add(pk2 a, pk2 b) {
    pk2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 1/x by the SFU without the denormal pre-scaling that __fdividef / __frcp_rn add
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// max(0, min(1, a * b + c)) in one FMA-pipe instruction (FFMA.SAT); NaN -> +0
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

}  // namespace drb
