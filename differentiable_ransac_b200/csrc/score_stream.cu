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
#include "device_cfg.cuh"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "msac_records.cuh"
#include "sampson.cuh"
#include "tile_pipe.cuh"

namespace drb {

constexpr int kSsLanes = 32;       // models per block (one warp)
constexpr int kSsPiece = 96;       // records (2 correspondences each) per work unit = one bulk copy of 3 KB
constexpr int kSsRing = 2;         // pieces in flight per warp (6 KB, as the one-CTA-per-block kernel)
constexpr int kSsSmemFloats = kSsRing * kSsPiece * 8;

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
        const float nci = -1.f / (tt * tt);

        float4* cf4 = ring_f4 + slot * kSsPiece * 2;
        mbar_wait(&bars[slot], (bar_phase >> slot) & 1u);
        bar_phase ^= 1u << slot;
        const int pts = min(2 * kSsPiece, N - 2 * cur.j * kSsPiece);
        const int recs_even = msac_interleave(cf4, pts, lane, kSsLanes);   // even capacity 96 > recs when padded
        __syncwarp();
        pk2 acc = pk2_splat(0.f), accB = pk2_splat(0.f);
        const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(cf4);
        for (int i = 0; i < recs_even; i += 2) msac_two_records(t2 + 2 * i, mp, nci, acc, accB);
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
    // CTAs per SM of this kernel x SMs of the CURRENT device, cached per device
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int grid = cache[dev].load(std::memory_order_relaxed);
    if (grid > 0) return grid;
    int per_sm = 16;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_msac_stream_kernel, kSsLanes, 0) != cudaSuccess ||
        per_sm < 1)
        per_sm = 16;
    grid = sm_count_current_device() * per_sm;
    cache[dev].store(grid, std::memory_order_relaxed);
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
