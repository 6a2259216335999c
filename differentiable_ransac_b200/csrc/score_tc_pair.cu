// Soft-MSAC scoring on the tensor cores, TWO SMs per tile (tcgen05 cta_group::2; ops.score_msac(kernel="tc_bf16p2")).
//
// Same contraction, operand images, epilogue arithmetic and C-ABI contract as score_tc.cu (scorings/msac_score.py:12-55,
// ransac.py:114), with the tile doubled across an SM pair.  Why: in score_tc.cu the tensor side
// -- not the hand-over, not the XU or FMA pipes -- is the sensitive part: every CTA reads 72 KB of operands from its
// shared memory per (128 x 128)-pair tile, 48 KB of them the model operand that does not change within a unit, and
// streams the pair's 384 KB of correspondence images from L2 once per unit (438 MB per launch).  With cta_group::2 a
// cluster of two CTAs shares a unit of 128 models:
//   * a tile is 256 correspondences: CTA r loads the image tile 2 t + r (its 128 rows of the A operand) -- each CTA
//     streams HALF of the pair's images per unit;
//   * the model operand (256 rows: an r and a j row per model) is split: CTA r builds and holds rows 128 r .. 128 r + 127
//     (models 64 r .. 64 r + 63) -- half the builder work and half the model-operand reads per CTA;
//   * the leader CTA (rank 0) issues six 256 x 256 x 16 MMAs per tile; the accumulator rows 128 r .. of a tile land in
//     CTA r's tensor memory, and each CTA's epilogue warps reduce THEIR 128 correspondences exactly as in score_tc.cu;
//   * at the end of a unit CTA 1 hands its 128 partial sums to CTA 0 through distributed shared memory, which adds them
//     in a fixed order (scores stay schedule-independent) and emits scores / arg-max keys.
// Cross-CTA signalling: the leader's MMA thread needs "both A stages full", "both B halves built", "both accumulators
// drained": the peer's otherwise idle MMA warp relays its a_full to a leader-side barrier, the peer's builders and
// epilogue warps arrive on the leader's b_full / d_empty remotely (mapa + mbarrier.arrive.release.cluster); the other
// direction (stage free, accumulator full, operand free) is one multicast tcgen05.commit to both CTAs.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "device_cfg.cuh"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "msac_tc_layout.cuh"
#include "sampson.cuh"
#include "tc_ptx.cuh"
#include "tile_pipe.cuh"

#ifndef DRB_TCP_EARLY_RELEASE
#define DRB_TCP_EARLY_RELEASE 0      // measured: 0.169 ms with, 0.165 ms without
#endif
#ifndef DRB_TCP_ABLATE               // profiles/microbench/tc_ablate.py: bit 0 no MMAs issued, bit 1 no epilogue arithmetic
#define DRB_TCP_ABLATE 0
#endif

namespace drb {
namespace tc {
int launch_features(bool bf16, const float* matches, int B, int N, int tiles, uint32_t* images, cudaStream_t s);   // score_tc.cu
}
namespace tcp {
using namespace drb::tc;

constexpr int kStages = 4;
constexpr int kWarpProducer = 0, kWarpMma = 1, kWarpBuild0 = 2, kBuildThreads = 64, kWarpEpi0 = 4, kEpi = 8;
constexpr int kThreads = (kWarpEpi0 + kEpi) * 32;          // 384
constexpr int kTmemCols = 512;
constexpr int kMaxPairs = 1024;
constexpr int kBHalfBytes = kBBytes / 2;                    // 24 576: this CTA's 128 rows of the model operand
constexpr int kHalfModels = kTileModels / 2;                // 64

constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStages * kABytes;            //  98 304
constexpr int kOffBars = kOffB + 2 * kBHalfBytes;           // 147 456
// a_full[4] a_peer[4] a_empty[4] d_full[2] d_empty[2] b_full[2] b_empty[2] unit[2]
constexpr int kNumBars = 3 * kStages + 2 + 2 + 2 + 2 + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kOffPrefix = kOffTmemPtr + 16;
constexpr int kOffPart = kOffPrefix + (kMaxPairs + 1) * 4 + 12;
constexpr int kOffPeer = kOffPart + 2 * 4 * kTileModels * 4;     // peer_part[unit parity][model]: written by CTA 1 into CTA 0
constexpr int kSmemBytes = kOffPeer + 2 * kTileModels * 4;
static_assert(kOffPart % 16 == 0 && kOffTmemPtr % 16 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// ---- cluster / cta_group::2 PTX --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* smem_ptr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the same without the cluster-scope release (a MEMBAR.GPU in front of every arrival: the first runs had it on the
// per-tile d_empty arrivals of the peer's eight epilogue warps, and the bare pipeline -- no MMAs, no arithmetic -- took
// 0.118 ms).  For arrivals that order no generic-proxy data: "my stage is full" (written by the bulk-copy engine),
// "my half of the operand is built" (fence.proxy.async before it), "I have drained the accumulator" (tcgen05 fence).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_remote_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// wait on a LOCAL barrier whose arrivals may come from the peer CTA: acquire at cluster scope; `sleep` for the idle roles
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, bool sleep) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        if (sleep) __nanosleep(100);
    }
}
// The other waits: the arrivals come from tcgen05.commit (stage free, accumulator full, operand free) or announce data
// that only the tensor core's async proxy or tcgen05.ld will touch (peer stage full, operand built, accumulator
// drained), so there is no generic-proxy memory to acquire and the default CTA scope is enough -- the cluster-scope
// acquire above costs a MEMBAR.GPU + CCTL.IVALL per successful poll (the first run with it everywhere: 0.169 ms).
__device__ __forceinline__ void mbar_wait_local(uint64_t* bar, uint32_t parity, bool sleep) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        if (sleep) __nanosleep(100);
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (BF16) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// one arrival on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void mma_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
score_msac_tc_pair_kernel(const uint32_t* __restrict__ images, const float* __restrict__ models,
                          const int32_t* __restrict__ count, const int32_t* __restrict__ ids, const float* __restrict__ thr,
                          int B, int M, int N, int tiles, float* __restrict__ scores,
                          unsigned long long* __restrict__ best_packed) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* a_full = bars;                      // this CTA's stage is in shared memory (bulk copy complete)
    uint64_t* a_peer = a_full + kStages;          // leader only: the PEER's stage is in its shared memory (relayed)
    uint64_t* a_empty = a_peer + kStages;         // both: the MMAs have read the stage (multicast commit)
    uint64_t* d_full = a_empty + kStages;         // both: the accumulator is complete (multicast commit)
    uint64_t* d_empty = d_full + 2;               // leader only: both CTAs' epilogue warps have drained it (2 * kEpi arrivals)
    uint64_t* b_full = d_empty + 2;               // leader only: both halves of the model operand are built (4 arrivals)
    uint64_t* b_empty = b_full + 2;               // both: the MMAs have read the model operand (multicast commit)
    uint64_t* unit_bar = b_empty + 2;             // leader only: the peer's partial sums of a unit have arrived (4 arrivals)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
    int* prefix = reinterpret_cast<int*>(smem + kOffPrefix);
    float* part = reinterpret_cast<float*>(smem + kOffPart);
    float* peer_part = reinterpret_cast<float*>(smem + kOffPeer);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0: leader
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int dtiles = (tiles + 1) >> 1;              // tiles of 256 correspondences

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_peer[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 2 * kEpi);
            mbar_init(&b_full[i], 2 * (kBuildThreads / 32));
            mbar_init(&b_empty[i], 1);
            mbar_init(&unit_bar[i], 4);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWarpEpi0) unit_prefix(prefix, count, B, M, kTileModels, lane);
    if (warp == kWarpMma) tmem_alloc2(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();           // the peer's barriers are initialised and its tensor memory allocated before anyone signals
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_units = prefix[B];

    if (warp == kWarpProducer) {
        // ===== producer: THIS CTA's half of every tile (image tile 2 t + rank) =====
        if (lane == 0) {
            Ring ra;
#pragma unroll 1
            for (int u = cluster_id; u < n_units; u += n_clusters) {
                int b, mt;
                unit_of(prefix, B, u, b, mt);
                const uint32_t* src = images + (size_t)b * (2 * dtiles) * (kABytes / 4);
#pragma unroll 1
                for (int t = 0; t < dtiles; ++t) {
                    mbar_wait_local(&a_empty[ra.idx], ra.phase ^ 1u, true);
                    mbar_expect_tx(&a_full[ra.idx], kABytes);
                    bulk_g2s(smem + kOffA + ra.idx * kABytes, src + (size_t)(2 * t + (int)rank) * (kABytes / 4), kABytes,
                             &a_full[ra.idx]);
                    ra.advance(kStages);
                }
            }
        }
    } else if (warp == kWarpMma) {
        if (lane == 0) {
            Ring ra, rd, rb;
            if (rank == 0) {
                // ===== MMA issuer (leader) =====
                const uint32_t idesc = instr_desc_mn(2 * kTileM, kTileN, BF16);
#pragma unroll 1
                for (int u = cluster_id; u < n_units; u += n_clusters) {
                    mbar_wait_local(&b_full[rb.idx], rb.phase, true);
                    const uint64_t bdesc = smem_desc(smem_u32(smem + kOffB + rb.idx * kBHalfBytes));
#pragma unroll 1
                    for (int t = 0; t < dtiles; ++t) {
                        mbar_wait_local(&d_empty[rd.idx], rd.phase ^ 1u, false);
                        mbar_wait_local(&a_full[ra.idx], ra.phase, false);
                        mbar_wait_local(&a_peer[ra.idx], ra.phase, false);
                        tc_fence_after();
                        const uint64_t adesc = smem_desc(smem_u32(smem + kOffA + ra.idx * kABytes));
                        const uint32_t d = tmem_base + (uint32_t)(rd.idx * kTileN);
                        DRB_UNROLL
                        for (int k = 0; k < ((DRB_TCP_ABLATE & 1) ? 0 : kKSteps); ++k)
                            mma2<BF16>(d, smem_desc_kstep(adesc, k), smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                        mma_commit2(&a_empty[ra.idx]);
                        mma_commit2(&d_full[rd.idx]);
                        ra.advance(kStages);
                        rd.advance(2);
                    }
                    mma_commit2(&b_empty[rb.idx]);
                    rb.advance(2);
                }
            } else {
                // ===== relay (peer): "my stage is full" -> the leader's a_peer =====
#pragma unroll 1
                for (int u = cluster_id; u < n_units; u += n_clusters) {
#pragma unroll 1
                    for (int t = 0; t < dtiles; ++t) {
                        mbar_wait_local(&a_full[ra.idx], ra.phase, true);
                        mbar_arrive_remote_relaxed(map_to_cta(&a_peer[ra.idx], 0));
                        ra.advance(kStages);
                    }
                }
            }
        }
    } else if (warp < kWarpEpi0) {
        // ===== builders: THIS CTA's half of the model operand (models 64 rank .. 64 rank + 63 of the unit) =====
        const int bt = threadIdx.x - kWarpBuild0 * 32;   // 0 .. 63
        Ring rb;
#pragma unroll 1
        for (int u = cluster_id; u < n_units; u += n_clusters) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            mbar_wait_local(&b_empty[rb.idx], rb.phase ^ 1u, true);
            uint32_t* img = reinterpret_cast<uint32_t*>(smem + kOffB + rb.idx * kBHalfBytes);
            const int i = (int)rank * kHalfModels + bt;   // model of the unit; its columns 4 (i >> 1) + .. are rows of MY half
            const int mi = mt * kTileModels + i;
            float m[9];
            DRB_UNROLL
            for (int q = 0; q < 9; ++q) m[q] = mi < cnt ? __ldg(models + ((size_t)b * M + mi) * 9 + q) : 0.f;
            float cr[kFeat], cj[kFeat];
            uint32_t row48[kK];
            model_rows(m, mi < cnt, true, cr, cj);
            const int row0 = (int)rank * (kTileN / 2);    // first D column (= operand row) of my half
            operand_row_words(cr, false, BF16, row48);
            DRB_UNROLL
            for (int c = 0; c < kK / 4; ++c)
                *reinterpret_cast<uint4*>(img + image_index(column_r(i) - row0, 4 * c)) =
                    make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
            operand_row_words(cj, false, BF16, row48);
            DRB_UNROLL
            for (int c = 0; c < kK / 4; ++c)
                *reinterpret_cast<uint4*>(img + image_index(column_j_swapped(i) - row0, 4 * c)) =
                    make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&b_full[rb.idx]);
                else mbar_arrive_remote_relaxed(map_to_cta(&b_full[rb.idx], 0));
            }
            rb.advance(2);
        }
    } else {
        // ===== epilogue: my 128 correspondences of every tile against the unit's 128 models (as in score_tc.cu) =====
        constexpr int kParts = kEpi / 4, kCols = kTileN / kParts, kChunks = kCols / 32, kAcc = kChunks * 8;
        const int et = threadIdx.x - kWarpEpi0 * 32;
        const int quarter = warp & 3;
        const int half = (warp - kWarpEpi0) >> 2;
        Ring rd;
        int parity = 0;
        uint32_t unit_phase[2] = {0u, 0u};
#pragma unroll 1
        for (int u = cluster_id; u < n_units; u += n_clusters, parity ^= 1) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            const float th = 1.5f * __ldg(thr + b);
            const float nci = -1.f / (th * th);
            pk2 acc[kAcc];
            DRB_UNROLL
            for (int i = 0; i < kAcc; ++i) acc[i] = pk2_splat(0.f);
#pragma unroll 1
            for (int t = 0; t < dtiles; ++t) {
                mbar_wait_local(&d_full[rd.idx], rd.phase, false);
                __syncwarp();
                tc_fence_after();
                const float one = ((2 * t + (int)rank) * kTileM + quarter * 32 + lane < N) ? 1.f : 0.f;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(rd.idx * kTileN + half * kCols);
                DRB_UNROLL
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t v[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
#if DRB_TCP_EARLY_RELEASE
                    if (c == kChunks - 1) {
                        // the accumulator is in registers: hand it back before the arithmetic of the last chunk -- the
                        // cross-CTA hand-over is slow enough (a remote arrive per warp) that starting it early pays here
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank == 0) mbar_arrive(&d_empty[rd.idx]);
                            else mbar_arrive_remote_relaxed(map_to_cta(&d_empty[rd.idx], 0));
                        }
                    }
#endif
                    if (DRB_TCP_ABLATE & 2) {
                        acc[c * 8] = pk2_add(acc[c * 8], pk2_make(__uint_as_float(v[0] ^ v[13]), __uint_as_float(v[31] ^ v[20])));
                        continue;
                    }
                    DRB_UNROLL
                    for (int q = 0; q < 8; ++q) {
                        const pk2 R = pk2_make(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
                        const pk2 R2 = pk2_mul(R, R);
                        const float ja = __uint_as_float(v[4 * q + 2]), jb = __uint_as_float(v[4 * q + 3]);
                        const float tn = rcp_approx(ja * jb) * nci;
                        float w0, w1;
                        pk2_split(pk2_mul(R2, pk2_make(ja, jb)), w0, w1);
                        acc[c * 8 + q] = pk2_add(acc[c * 8 + q], pk2_make(fma_sat(w0, tn, one), fma_sat(w1, tn, one)));
                    }
                }
#if !DRB_TCP_EARLY_RELEASE
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(&d_empty[rd.idx]);
                    else mbar_arrive_remote_relaxed(map_to_cta(&d_empty[rd.idx], 0));
                }
#endif
                rd.advance(2);
            }
            // ---- my 128 rows: butterfly over the lanes, then the four quarters (score_tc.cu) ---------------------
            float a[2 * kAcc];
            DRB_UNROLL
            for (int i = 0; i < kAcc; ++i) pk2_split(acc[i], a[2 * i], a[2 * i + 1]);
            DRB_UNROLL
            for (int w = kAcc, o = 16; o > 0; w >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
                DRB_UNROLL
                for (int i = 0; i < w; ++i) {
                    const float keep = up ? a[w + i] : a[i];
                    const float send = up ? a[i] : a[w + i];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            constexpr int kPer = kAcc / 16;
            float* pp = part + (size_t)parity * (4 * kTileModels);
            DRB_UNROLL
            for (int i = 0; i < kPer; ++i) pp[quarter * kTileModels + half * (kCols / 2) + kPer * lane + i] = a[i];
            asm volatile("bar.sync 1, %0;" ::"n"(kEpi * 32) : "memory");
            if (et < kTileModels) {
                const float mine = ((pp[0 * kTileModels + et] + pp[1 * kTileModels + et]) + pp[2 * kTileModels + et]) +
                                   pp[3 * kTileModels + et];
                if (rank != 0) {
                    // hand my partial sum to the leader: distributed shared memory, then one arrival per warp
                    st_remote_f32(map_to_cta(&peer_part[parity * kTileModels + et], 0), mine);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(map_to_cta(&unit_bar[parity], 0));
                } else {
                    mbar_wait_cluster(&unit_bar[parity], unit_phase[parity], false);
                    unit_phase[parity] ^= 1u;
                    const float score = mine + peer_part[parity * kTileModels + et];     // rows of CTA 0, then rows of CTA 1
                    const int mi = mt * kTileModels + et;
                    const bool live = mi < cnt;
                    if (live && scores) scores[(size_t)b * M + mi] = score;
                    unsigned long long key = live ? pack_best(score, ids ? __ldg(ids + (size_t)b * M + mi) : mi) : 0ull;
                    DRB_UNROLL
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                        key = other > key ? other : key;
                    }
                    if (lane == 0 && key) atomicMax(best_packed + b, key);
                }
            }
        }
    }

    // ---- teardown: both CTAs are done with both tensor memories before either frees ---------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kWarpMma) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, kTmemCols);
    }
}

template <bool BF16>
static int launch(const float* matches, const float* models, const int32_t* count, const int32_t* ids, const float* thr,
                  int B, int M, int N, float* scores, unsigned long long* best_packed, uint32_t* images, cudaStream_t s) {
    static std::atomic<unsigned long long> configured{0};
    if (!ensure_dynamic_smem(score_msac_tc_pair_kernel<BF16>, kSmemBytes, configured)) return DRB_ERR_CUDA;
    const int tiles = (N + kTileM - 1) / kTileM;
    const int rc = drb::tc::launch_features(BF16, matches, B, N, 2 * ((tiles + 1) / 2), images, s);   // an EVEN tile count per pair
    if (rc != DRB_OK) return rc;
    const long long max_units = (long long)B * ((M + kTileModels - 1) / kTileModels);
    const int sms = sm_count_current_device();
    int clusters = (int)(max_units < sms / 2 ? max_units : sms / 2);
    if (clusters < 1) clusters = 1;
    score_msac_tc_pair_kernel<BF16><<<2 * clusters, kThreads, kSmemBytes, s>>>(images, models, count, ids, thr, B, M, N, tiles,
                                                                              scores, best_packed);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

int dispatch(bool bf16, const float* matches, const float* models, const int32_t* count, const int32_t* ids,
             const float* thr, int B, int M, int N, float* scores, unsigned long long* best_packed, uint32_t* images,
             cudaStream_t s) {
    return bf16 ? launch<true>(matches, models, count, ids, thr, B, M, N, scores, best_packed, images, s)
                : launch<false>(matches, models, count, ids, thr, B, M, N, scores, best_packed, images, s);
}

}  // namespace tcp
}  // namespace drb
