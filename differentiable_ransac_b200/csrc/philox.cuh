// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the
// uniform -> Gumbel(0,1) transform shared by the sampler forward and backward
// (the backward regenerates the forward's noise instead of storing K x N floats).
#pragma once

#include "drb_common.cuh"

namespace drb {

struct Philox4 {
    uint32_t x, y, z, w;
};

DRB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    DRB_UNROLL
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
        uint32_t lo0, hi0, lo1, hi1;
        asm("{.reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo0), "=r"(hi0) : "r"(M0), "r"(c0));
        asm("{.reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo1), "=r"(hi1) : "r"(M1), "r"(c2));
#else
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t lo0 = (uint32_t)p0, hi0 = (uint32_t)(p0 >> 32), lo1 = (uint32_t)p1, hi1 = (uint32_t)(p1 >> 32);
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 23 random bits -> v in [2^-24, 1 - 2^-24] (both ends exactly representable), then
// G = -log(-log v).  The clamp keeps the double logarithm finite when the fast
// log2 rounds -log v to <= 0 next to v = 1.
// Built in the mantissa instead of converted: 1.m - (1 - 2^-24) = (m + 0.5) 2^-23 exactly, one FADD on the FMA
// pipe where int -> float would be a quarter-rate op on the same unit as the logarithms that follow.
DRB_D float uniform_from_bits(uint32_t r) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(0x3f800000u | (r >> 9)) - 0.99999994039535522f;
#else
    return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-07f;
#endif
}

#if defined(__CUDACC__)
// raw SFU log2: every argument here is a normal number, so the denormal pre-scaling that
// __logf / __log2f wrap around MUFU.LG2 is dead weight
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif

DRB_D float gumbel_from_bits(uint32_t r) {
    const float v = uniform_from_bits(r);
#if defined(__CUDA_ARCH__)
    const float e = fmaxf(-0.6931471805599453f * lg2_approx(v), 1e-10f);
    return -0.6931471805599453f * lg2_approx(e);
#else
    const float e = fmaxf(-logf(v), 1e-10f);
    return -logf(e);
#endif
}

}  // namespace drb
