// Soft-MSAC scoring on a work queue: persistent one-warp CTAs pull (model block, piece of the
// correspondences) units from an atomic counter.
//
// Replaces scorings/msac_score.py:12-55 + the arg-max of ransac.py:114 -- same contract as
// drb_score_msac (score.cu), plus a caller-provided workspace.
//
// Why (ncu, profiles/r1_notes.md): with one CTA per (pair, 32 models) all ~4 300 live warps of the headline
// shape start together, but the warp scheduler is not fair -- it keeps favouring the same warps of a
// sub-partition, so they finish one after the other: a warp lives for 0.58 of the kernel on average, 4.7
// of the 7.3 warps of a sub-partition are alive on average, and once fewer than ~4 are left they cannot keep
// the FMA pipe busy (71 % busy over the kernel).  Cutting the same static work into equal shares changes
// nothing.  Here a warp that finishes early simply takes the next unit, so every sub-partition keeps all
// its warps until the queue is empty; the tail is one piece (96 records) instead of half the kernel.
//
// A record is two correspondences, interleaved in shared memory for the packed-fp32 inner loop
// (x1p x1q y1p y1q | x2p x2q y2p y2q); correspondences past the end are NaN: max(NaN, 0) = 0, so odd
// tails need no scalar path.  The pieces of a block are summed in piece order by whichever warp delivers
// the last one (scores do not depend on the schedule); the arg-max is fused as before.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "sampson.cuh"
#include "tile_pipe.cuh"

namespace drb {

constexpr int kSsLanes = 32;       // models per block (one warp)
constexpr int kSsPiece = 96;       // records (2 correspondences each) per work unit = one bulk copy of 3 KB
constexpr int kSsRing = 2;         // pieces in flight per warp (6 KB, as the one-CTA-per-block kernel)
constexpr int kSsSmemFloats = kSsRing * kSsPiece * 8;

// Two records (A, B = four correspondences) against the thread's model: packed Sampson residuals and the
// soft-MSAC increment.  Same hand-grouped order as score_msac_kernel<true> (operands shared by neighbours).
__device__ __forceinline__ void ss_two_records(const ulonglong2* t2, const pk2* mp, pk2 neg_inv, pk2 one, pk2& acc,
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
    float aul, auh, bul, buh, nci, dummy;
    pk2_split(AU, aul, auh);
    pk2_split(BU, bul, buh);
    pk2_split(neg_inv, nci, dummy);
    acc = pk2_add(acc, pk2_make(fma_sat(aul, nci, 1.f), fma_sat(auh, nci, 1.f)));
    accB = pk2_add(accB, pk2_make(fma_sat(bul, nci, 1.f), fma_sat(buh, nci, 1.f)));
}

__device__ __forceinline__ int ss_count(const int32_t* count, int b, int M) {
    return count ? min(max(__ldg(count + b), 0), M) : M;
}
__device__ __forceinline__ int ss_blocks(const int32_t* count, int b, int M) {
    return (ss_count(count, b, M) + kSsLanes - 1) / kSsLanes;
}

struct SsUnit {
    int b, mb, j;   // pair, model block within the pair, piece; b < 0: the queue is empty
};

__global__ void __launch_bounds__(kSsLanes, 32)
score_msac_stream_kernel(const float* __restrict__ matches, const float* __restrict__ models,
                         const int32_t* __restrict__ count, const int32_t* __restrict__ ids,
                         const float* __restrict__ thr, int B, int M, int N, float* __restrict__ scores,
                         unsigned long long* __restrict__ best_packed, int32_t* __restrict__ ctrl,
                         int32_t* __restrict__ arrive, float* __restrict__ scratch) {
    __shared__ __align__(128) float smem[kSsSmemFloats];
    __shared__ __align__(8) uint64_t bars[kSsRing];
    const int lane = threadIdx.x;
    const int NP = (N + 1) >> 1;                          // records per model block
    const int pieces = (NP + kSsPiece - 1) / kSsPiece;    // work units per model block
    const int mblocks = (M + kSsLanes - 1) / kSsLanes;    // arrive[] / scratch stride per pair

    // Each lane owns a contiguous range of R pairs; start = number of model blocks before its range.
    const int R = (B + 31) / 32;
    int mine = 0;
    for (int b = lane * R; b < min(B, (lane + 1) * R); ++b) mine += ss_blocks(count, b, M);
    int incl = mine;
    DRB_UNROLL
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int start = incl - mine;
    const int G = __shfl_sync(0xffffffffu, incl, 31);
    const long long total = (long long)G * pieces;

    auto unit_of = [&](long long u) {   // flattened index -> unit; every lane returns the same value
        SsUnit un;
        un.b = -1;
        un.mb = un.j = 0;
        if (u >= total) return un;
        const int g = (int)(u / pieces);
        un.j = (int)(u - (long long)g * pieces);
        const int owner = __popc(__ballot_sync(0xffffffffu, start <= g)) - 1;   // last lane whose range starts at or before g
        int b = -1, mb = 0;
        if (lane == owner) {
            int before = start;
            for (b = lane * R;; ++b) {
                const int nb = ss_blocks(count, b, M);
                if (g < before + nb) break;
                before += nb;
            }
            mb = g - before;
        }
        un.b = __shfl_sync(0xffffffffu, b, owner);
        un.mb = __shfl_sync(0xffffffffu, mb, owner);
        return un;
    };
    auto issue = [&](const SsUnit& un, int slot) {   // lane 0: bulk copy of the unit's correspondences
        const int p0 = un.j * kSsPiece;
        const int pts = min(2 * kSsPiece, N - 2 * p0);
        mbar_expect_tx(&bars[slot], (uint32_t)pts * 16u);
        bulk_g2s(smem + slot * kSsPiece * 8, matches + ((size_t)un.b * N + 2 * (size_t)p0) * 4, (uint32_t)pts * 16u,
                 &bars[slot]);
    };

    if (lane == 0) {
        DRB_UNROLL
        for (int i = 0; i < kSsRing; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncwarp();
    uint32_t bar_phase = 0;   // bit i = parity to wait for on bars[i]
    float4* ring_f4 = reinterpret_cast<float4*>(smem);
    const float nanf_ = __int_as_float(0x7fc00000);

    // The first unit of every warp is fixed (unit w: no 4 736 simultaneous first pops of one counter); later
    // ones come from the queue.  The next unit is popped, and its correspondences requested, before the
    // current one is processed.
    const long long W = gridDim.x;
    auto pop = [&]() {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(reinterpret_cast<unsigned long long*>(ctrl), 1ull);
        return unit_of((long long)__shfl_sync(0xffffffffu, ticket, 0) + W);
    };
    SsUnit cur = unit_of(blockIdx.x);
    if (cur.b >= 0 && lane == 0) issue(cur, 0);
    int slot = 0;
    while (cur.b >= 0) {
        const SsUnit nxt = pop();
        if (nxt.b >= 0 && lane == 0) {
            fence_proxy_async();   // the other slot was last read / written through the generic proxy
            issue(nxt, slot ^ 1);
        }
        const int b = cur.b, mb = cur.mb;
        const int cnt = ss_count(count, b, M);
        const int mi = mb * kSsLanes + lane;
        const bool live = mi < cnt;
        pk2 mp[9];
        {
            const float* src = models + ((size_t)b * M + (live ? mi : mb * kSsLanes)) * 9;
            DRB_UNROLL
            for (int i = 0; i < 9; ++i) mp[i] = pk2_splat(__ldg(src + i));
        }
        const float tt = 1.5f * __ldg(thr + b);
        const pk2 nc = pk2_splat(-1.f / (tt * tt)), one = pk2_splat(1.f);

        float4* cf4 = ring_f4 + slot * kSsPiece * 2;
        mbar_wait(&bars[slot], (bar_phase >> slot) & 1u);
        bar_phase ^= 1u << slot;
        // interleave in place: (x1p y1p x2p y2p | x1q y1q x2q y2q) -> (x1p x1q y1p y1q | x2p x2q y2p y2q);
        // a correspondence past the end becomes NaN (contributes exactly 0)
        const int pts = min(2 * kSsPiece, N - 2 * cur.j * kSsPiece);
        const int recs = (pts + 1) >> 1;
        DRB_UNROLL
        for (int k = 0; k < kSsPiece / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < recs) {
                const float4 p = cf4[2 * i];
                float4 q = cf4[2 * i + 1];
                if (2 * i + 1 >= pts) q = make_float4(nanf_, nanf_, nanf_, nanf_);
                cf4[2 * i] = make_float4(p.x, q.x, p.y, q.y);
                cf4[2 * i + 1] = make_float4(p.z, q.z, p.w, q.w);
            } else if (i == recs && (recs & 1)) {   // pad to an even number of records
                cf4[2 * i] = make_float4(nanf_, nanf_, nanf_, nanf_);
                cf4[2 * i + 1] = make_float4(nanf_, nanf_, nanf_, nanf_);
            }
        }
        __syncwarp();
        pk2 acc = pk2_splat(0.f), accB = pk2_splat(0.f);
        const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(cf4);
        const int recs_even = (recs + 1) & ~1;
        for (int i = 0; i < recs_even; i += 2) ss_two_records(t2 + 2 * i, mp, nc, one, acc, accB);
        __syncwarp();   // every lane is done with this slot before lane 0 refills it
        float lo, hi;
        pk2_split(pk2_add(acc, accB), lo, hi);
        float score = lo + hi;

        // ---- combine with the other pieces of this block (if any), arg-max -------------------------
        bool finish = true;
        if (pieces > 1) {
            const size_t blk = (size_t)b * mblocks + mb;
            scratch[(blk * pieces + cur.j) * kSsLanes + lane] = score;
            __threadfence();
            __syncwarp();
            int old = 0;
            if (lane == 0) old = atomicAdd(arrive + blk, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            finish = old == pieces - 1;
            if (finish) {
                __threadfence();
                score = 0.f;
                for (int j = 0; j < pieces; ++j)   // piece order: the sum does not depend on the schedule
                    score += __ldcg(scratch + (blk * pieces + j) * kSsLanes + lane);
                if (lane == 0) arrive[blk] = 0;   // leave the workspace zeroed for the next call
            }
        }
        if (finish) {
            if (live && scores) scores[(size_t)b * M + mi] = score;
            unsigned long long key = live ? pack_best(score, ids ? __ldg(ids + (size_t)b * M + mi) : mi) : 0ull;
            DRB_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) atomicMax(best_packed + b, key);
        }
        cur = nxt;
        slot ^= 1;
    }
    // the last warp to leave rewinds the queue (the workspace stays zeroed between calls)
    if (lane == 0) {
        __threadfence();
        const int gone = atomicAdd(ctrl + 2, 1);
        if (gone == (int)gridDim.x - 1) {
            ctrl[0] = 0;
            ctrl[1] = 0;
            ctrl[2] = 0;
        }
    }
}

static int ss_grid() {
    static const int grid = []() {
        int dev = 0, sms = 148, per_sm = 16;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_msac_stream_kernel, kSsLanes, 0) != cudaSuccess ||
            per_sm < 1)
            per_sm = 16;
        return sms * per_sm;
    }();
    return grid;
}

// workspace layout: [ctrl: 4 x int32 | arrive: B * mblocks x int32] (zero on entry, left zero) | scratch
static size_t ss_zeroed_bytes(int B, int M) {
    return ((4 + (size_t)B * ((M + kSsLanes - 1) / kSsLanes)) * sizeof(int32_t) + 255) & ~(size_t)255;
}

}  // namespace drb

using namespace drb;

extern "C" size_t drb_score_msac_workspace_bytes(int B, int M, int N) {
    if (B <= 0 || M <= 0 || N <= 0) return 0;
    const size_t pieces = (((size_t)N + 1) / 2 + kSsPiece - 1) / kSsPiece;
    const size_t scratch = pieces > 1 ? (size_t)B * ((M + kSsLanes - 1) / kSsLanes) * pieces * kSsLanes * sizeof(float) : 0;
    return ss_zeroed_bytes(B, M) + scratch;
}

extern "C" size_t drb_score_msac_workspace_zeroed_bytes(int B, int M) {
    return (B <= 0 || M <= 0) ? 0 : ss_zeroed_bytes(B, M);
}

extern "C" int drb_score_msac_stream(const float* matches, const float* models, const int32_t* count,
                                     const int32_t* ids, const float* thr, int B, int M, int N, float* scores,
                                     unsigned long long* best_packed, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!matches || !models || !thr || !best_packed || !workspace) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || B > 1024 || M <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    if (workspace_bytes < drb_score_msac_workspace_bytes(B, M, N) || (reinterpret_cast<uintptr_t>(workspace) & 15))
        return DRB_ERR_BAD_SHAPE;
    int32_t* ctrl = reinterpret_cast<int32_t*>(workspace);
    float* scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ss_zeroed_bytes(B, M));
    score_msac_stream_kernel<<<ss_grid(), kSsLanes, 0, (cudaStream_t)stream>>>(matches, models, count, ids, thr, B, M, N,
                                                                               scores, best_packed, ctrl, ctrl + 4,
                                                                               scratch);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
