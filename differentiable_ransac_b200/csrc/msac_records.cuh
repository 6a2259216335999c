// Shared by the two soft-MSAC kernels (score.cu: one CTA per 32 models; score_stream.cu: work queue): the
// in-place interleave of a tile of correspondences into two-correspondence records and the packed-fp32
// Sampson / soft-MSAC arithmetic of two records against the thread's model.
//
// Replaces the arithmetic of scorings/msac_score.py:26-52: d2 = (x2' M x1)^2 / ((M x1)_0^2 + (M x1)_1^2 +
// (M' x2)_0^2 + (M' x2)_1^2), score += max(0, 1 - d2 / (1.5 thr)^2).
#pragma once

#include <cuda_runtime.h>

#include "f32x2.cuh"

namespace drb {

// Interleave `pts` correspondences (16 bytes each, at cf4) in place into (pts + 1) / 2 records
//   (x1p y1p x2p y2p | x1q y1q x2q y2q) -> (x1p x1q y1p y1q | x2p x2q y2p y2q)
// and pad to an EVEN number of records.  A correspondence that does not exist becomes NaN: every packed lane
// it touches ends in max(0, min(1, NaN)) = 0 (FFMA.SAT), so tails need no scalar path.  The tile must have
// room for the padding record (it has whenever pts is below the tile's even capacity).  Returns the number
// of records to process (even).  Call with all `nthreads` threads that share the tile, then synchronise them.
__device__ __forceinline__ int msac_interleave(float4* cf4, int pts, int tid, int nthreads) {
    const float nanf_ = __int_as_float(0x7fc00000);
    const int recs = (pts + 1) >> 1;
    const int recs_even = (recs + 1) & ~1;
    for (int i = tid; i < recs_even; i += nthreads) {
        if (i < recs) {
            const float4 p = cf4[2 * i];
            float4 q = cf4[2 * i + 1];
            if (2 * i + 1 >= pts) q = make_float4(nanf_, nanf_, nanf_, nanf_);
            cf4[2 * i] = make_float4(p.x, q.x, p.y, q.y);
            cf4[2 * i + 1] = make_float4(p.z, q.z, p.w, q.w);
        } else {
            cf4[2 * i] = make_float4(nanf_, nanf_, nanf_, nanf_);
            cf4[2 * i + 1] = make_float4(nanf_, nanf_, nanf_, nanf_);
        }
    }
    return recs_even;
}

// Two records (A, B = four correspondences) against the thread's model: packed Sampson residuals and the
// soft-MSAC increment; mp[i] = the model coefficient i in both lanes, neg_inv_thr2 = -1 / (1.5 thr)^2.  The packed
// instructions are written in an order in which neighbours share a source operand (the operand-reuse cache only
// serves adjacent instructions; ptxas still reschedules, profiles/r1_notes.md).
__device__ __forceinline__ void msac_two_records(const ulonglong2* t2, const pk2* mp, float neg_inv_thr2, pk2& acc,
                                                 pk2& accB) {
    const ulonglong2 a1 = t2[0], a2 = t2[1], b1 = t2[2], b2 = t2[3];
    const pk2 AX1 = a1.x, AY1 = a1.y, AX2 = a2.x, AY2 = a2.y;
    const pk2 BX1 = b1.x, BY1 = b1.y, BX2 = b2.x, BY2 = b2.y;
    const pk2 At0 = pk2_fma_v(mp[1], AY1, mp[2]);
    const pk2 At1 = pk2_fma_v(mp[4], AY1, mp[5]);
    const pk2 At2 = pk2_fma_v(mp[7], AY1, mp[8]);
    const pk2 Bt0 = pk2_fma_v(mp[1], BY1, mp[2]);
    const pk2 Bt1 = pk2_fma_v(mp[4], BY1, mp[5]);
    const pk2 Bt2 = pk2_fma_v(mp[7], BY1, mp[8]);
    const pk2 AE0 = pk2_fma_v(mp[0], AX1, At0);
    const pk2 AE1 = pk2_fma_v(mp[3], AX1, At1);
    const pk2 AE2 = pk2_fma_v(mp[6], AX1, At2);
    const pk2 BE0 = pk2_fma_v(mp[0], BX1, Bt0);
    const pk2 BE1 = pk2_fma_v(mp[3], BX1, Bt1);
    const pk2 BE2 = pk2_fma_v(mp[6], BX1, Bt2);
    const pk2 Au0 = pk2_fma_v(mp[3], AY2, mp[6]);
    const pk2 Au1 = pk2_fma_v(mp[4], AY2, mp[7]);
    const pk2 Bu0 = pk2_fma_v(mp[3], BY2, mp[6]);
    const pk2 Bu1 = pk2_fma_v(mp[4], BY2, mp[7]);
    const pk2 AF0 = pk2_fma_v(mp[0], AX2, Au0);
    const pk2 AF1 = pk2_fma_v(mp[1], AX2, Au1);
    const pk2 BF0 = pk2_fma_v(mp[0], BX2, Bu0);
    const pk2 BF1 = pk2_fma_v(mp[1], BX2, Bu1);
    const pk2 Ar0 = pk2_fma_v(AY2, AE1, AE2);
    const pk2 Br0 = pk2_fma_v(BY2, BE1, BE2);
    const pk2 Aj0 = pk2_mul_v(AF1, AF1);
    const pk2 Bj0 = pk2_mul_v(BF1, BF1);
    const pk2 AR = pk2_fma_v(AX2, AE0, Ar0);
    const pk2 BR = pk2_fma_v(BX2, BE0, Br0);
    const pk2 Aj1 = pk2_fma_v(AF0, AF0, Aj0);
    const pk2 Bj1 = pk2_fma_v(BF0, BF0, Bj0);
    const pk2 Aj2 = pk2_fma_v(AE1, AE1, Aj1);
    const pk2 Bj2 = pk2_fma_v(BE1, BE1, Bj1);
    const pk2 AJ = pk2_fma_v(AE0, AE0, Aj2);
    const pk2 BJ = pk2_fma_v(BE0, BE0, Bj2);
    const pk2 AR2 = pk2_mul_v(AR, AR);
    const pk2 BR2 = pk2_mul_v(BR, BR);
    float ajl, ajh, bjl, bjh;
    pk2_split(AJ, ajl, ajh);
    pk2_split(BJ, bjl, bjh);
    const pk2 AU = pk2_mul(AR2, pk2_make(rcp_approx(ajl), rcp_approx(ajh)));
    const pk2 BU = pk2_mul(BR2, pk2_make(rcp_approx(bjl), rcp_approx(bjh)));
    // 1 - u / thr^2 <= 1 always, so the saturating FMA is the clamp max(., 0) (and NaN -> 0)
    float aul, auh, bul, buh;
    const float nci = neg_inv_thr2;
    pk2_split(AU, aul, auh);
    pk2_split(BU, bul, buh);
    acc = pk2_add(acc, pk2_make(fma_sat(aul, nci, 1.f), fma_sat(auh, nci, 1.f)));
    accB = pk2_add(accB, pk2_make(fma_sat(bul, nci, 1.f), fma_sat(buh, nci, 1.f)));
}

}  // namespace drb
