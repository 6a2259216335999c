// Sampson residual pieces and the packed (score, id) arg-max key shared by the scoring kernels (score.cu)
// and the chunk bookkeeping of the adaptive driver (adaptive.cu).
#pragma once

#include <cuda_runtime.h>

namespace drb {

__device__ __forceinline__ unsigned long long pack_best(float score, int id) {
    // scores are >= 0, so the float bit pattern is monotone; ~id breaks ties towards the
    // lowest id (torch.argmax returns the first maximum).  NaN never wins.
    if (!(score >= 0.f)) return 0ull;
    return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)id);
}

// Sampson distance pieces for x2^T M x1
struct Sampson {
    float r, j;
};
__device__ __forceinline__ Sampson sampson(const float* m, float x1, float y1, float x2, float y2) {
    const float e0 = fmaf(m[0], x1, fmaf(m[1], y1, m[2]));
    const float e1 = fmaf(m[3], x1, fmaf(m[4], y1, m[5]));
    const float e2 = fmaf(m[6], x1, fmaf(m[7], y1, m[8]));
    const float f0 = fmaf(m[0], x2, fmaf(m[3], y2, m[6]));
    const float f1 = fmaf(m[1], x2, fmaf(m[4], y2, m[7]));
    Sampson s;
    s.r = fmaf(x2, e0, fmaf(y2, e1, e2));
    s.j = fmaf(e0, e0, fmaf(e1, e1, fmaf(f0, f0, f1 * f1)));
    return s;
}

// Decode the id of a packed key (-1 when nothing was ever packed).
__device__ __forceinline__ int packed_id(unsigned long long key) {
    return key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -1;
}
__device__ __forceinline__ float packed_score(unsigned long long key) {
    return key ? __uint_as_float((unsigned)(key >> 32)) : 0.f;
}

}  // namespace drb
