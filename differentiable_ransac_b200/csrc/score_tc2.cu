// Soft-MSAC scoring on the tensor cores, second arrangement: the MODELS stay in tensor memory, the
// correspondences stream through shared memory (opt-in: ops.score_msac(kernel="tc2_tf32" | "tc2_bf16" | ..._e16);
// measured on the B200 in round 2: oracle parity green, 0.128-0.146 ms at cfg2 against 0.1075 ms for score_tc.cu's
// pair variant -- the shared-memory operand traffic this arrangement saves was not the limiter).
//
// Same contraction and the same reference lines as score_tc.cu (scorings/msac_score.py:12-55, ransac.py:114;
// operands by msac_tc_layout.cuh), with the roles of the two operands swapped.  Why: in score_tc.cu both operands
// of every MMA come from shared memory (12 KB per MMA, 72 KB per tile of 128 x 128 pairs; ncu: the tensor core's
// shared-memory operand path at 33 % of its peak) and the model operand -- 48 KB, the larger one -- is read again
// for every tile of correspondences although it does not
// change within a unit.  Here a unit's 128 models are written ONCE into tensor memory (tcgen05.st from the
// registers of the warps that compute the coefficient words; no shared-memory image of the models exists) and
// serve as the A operand of every MMA of the unit; only the 15 KB tile of 80 correspondences is read from shared
// memory: 24 KB per 128 x 128 pairs instead of 72.  Two MMAs per K step (r and j have separate accumulators,
// D_r[model, point] and D_j[model, point]), so a thread of the epilogue -- lane = model -- finds r and j of the
// same pair in the same lane and sums over the points of its columns into ONE register pair: no butterfly, ~110
// (8 epilogue warps) or ~70 (16) registers per thread, 70 KB of shared memory -- which also leaves room on the SM
// for the 5-point CTAs of the next batch (score_tc.cu's CTA owns the whole register file).
//
// Tensor memory (512 columns): [0,160) accumulator 0 (r: 80 columns, j: 80), [160,320) accumulator 1,
// [320,416) model operand 0 (r: 48 columns, j: 48), [416,512) model operand 1.
// Warps: 0 bulk-copy producer, 1 TMEM allocation + MMA issue, 2-5 builders (one per lane quarter), 6.. epilogue.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "device_cfg.cuh"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "msac_tc_layout.cuh"
#include "sampson.cuh"
#include "tc_ptx.cuh"
#include "tile_pipe.cuh"

#ifndef DRB_TC_ABLATE   // profiles/microbench/tc_ablate.py (see score_tc.cu): bit 0 no MMAs, bit 1 no epilogue arithmetic
#define DRB_TC_ABLATE 0
#endif

namespace drb {
namespace tc2 {
using namespace drb::tc;

constexpr int kPts = 80;                          // correspondences per tile = MMA N
constexpr int kPtBytes = (kPts / 8) * kSBO;       // 15360
constexpr int kStages = 4;
constexpr int kModels = 128;                      // models per unit = MMA M = TMEM lanes
constexpr int kColD0 = 0, kColDStride = 2 * kPts; // accumulator b: r at kColD0 + b * 160, j at + 80
constexpr int kColA0 = 2 * kColDStride, kColAStride = 2 * kK;   // model operand b: r at 320 + b * 96, j at + 48
constexpr int kTmemCols = 512;
static_assert(kColA0 + 2 * kColAStride <= kTmemCols, "tensor memory budget");
constexpr int kWarpProducer = 0, kWarpMma = 1, kWarpBuild0 = 2, kBuildWarps = 4, kWarpEpi0 = kWarpBuild0 + kBuildWarps;
constexpr int threads_of(int epi_warps) { return (kWarpEpi0 + epi_warps) * 32; }
constexpr int kMaxPairs = 1024;

constexpr int kOffPts = 0;
constexpr int kOffBars = kOffPts + kStages * kPtBytes;            // 61440
constexpr int kNumBars = 2 * kStages + 2 + 2 + 2 + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kOffPrefix = kOffTmemPtr + 16;
constexpr int kOffPart = kOffPrefix + (kMaxPairs + 1) * 4 + 12;
constexpr int kSmemBytes = kOffPart + 2 * 4 * kModels * 4;        // part[unit parity][column part][model]
static_assert(kOffPart % 16 == 0, "alignment");

// ---- launch 1: correspondences -> operand images, tiles of kPts rows ------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(96)
msac_tc2_features_kernel(const float* __restrict__ matches, int N, int tiles, uint32_t* __restrict__ images) {
    const int b = blockIdx.y, t = blockIdx.x, row = threadIdx.x;
    if (row >= kPts) return;
    const int n = t * kPts + row;
    uint32_t row48[kK];
    if (n < N) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)b * N + n);
        float f[kFeat];
        features(p.x, p.y, p.z, p.w, f);
        operand_row_words(f, true, BF16, row48);
    } else {
        // a row past N: r = j = 0 -> 0 * rcp(0) = NaN -> FFMA.SAT -> 0
        DRB_UNROLL
        for (int k = 0; k < kK; ++k) row48[k] = 0u;
    }
    uint32_t* img = images + ((size_t)b * tiles + t) * (kPtBytes / 4);
    DRB_UNROLL
    for (int c = 0; c < kK / 4; ++c)
        *reinterpret_cast<uint4*>(img + image_index(row, 4 * c)) =
            make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
}

// ---- launch 2 -----------------------------------------------------------------------------------------
template <bool BF16, int EPI, bool PAIR>
__global__ void __launch_bounds__(threads_of(EPI), 1)
score_msac_tc2_kernel(const uint32_t* __restrict__ images, const float* __restrict__ models,
                      const int32_t* __restrict__ count, const int32_t* __restrict__ ids, const float* __restrict__ thr,
                      int B, int M, int N, int tiles, float* __restrict__ scores,
                      unsigned long long* __restrict__ best_packed) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* p_full = bars;                       // correspondence stages
    uint64_t* p_empty = bars + kStages;
    uint64_t* d_full = bars + 2 * kStages;         // accumulators
    uint64_t* d_empty = d_full + 2;
    uint64_t* a_full = d_empty + 2;                // model operands in tensor memory
    uint64_t* a_empty = a_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
    int* prefix = reinterpret_cast<int*>(smem + kOffPrefix);
    float* part = reinterpret_cast<float*>(smem + kOffPart);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&p_full[i], 1);
            mbar_init(&p_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], EPI);
            mbar_init(&a_full[i], kBuildWarps);
            mbar_init(&a_empty[i], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWarpEpi0) unit_prefix(prefix, count, B, M, kModels, lane);
    if (warp == kWarpMma) tmem_alloc(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_units = prefix[B];

    if (warp == kWarpProducer) {
        if (lane == 0) {
            Ring rp;
#pragma unroll 1
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int b, mt;
                unit_of(prefix, B, u, b, mt);
                const uint32_t* src = images + (size_t)b * tiles * (kPtBytes / 4);
#pragma unroll 1
                for (int t = 0; t < tiles; ++t) {
                    mbar_wait(&p_empty[rp.idx], rp.phase ^ 1u);
                    mbar_expect_tx(&p_full[rp.idx], kPtBytes);
                    bulk_g2s(smem + kOffPts + rp.idx * kPtBytes, src + (size_t)t * (kPtBytes / 4), kPtBytes, &p_full[rp.idx]);
                    rp.advance(kStages);
                }
            }
        }
    } else if (warp == kWarpMma) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_mn(kModels, kPts, BF16);
            Ring rp, rd, ra;
#pragma unroll 1
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                mbar_wait(&a_full[ra.idx], ra.phase);
                const uint32_t a_r = tmem_base + (uint32_t)(kColA0 + ra.idx * kColAStride);
                const uint32_t a_j = a_r + (uint32_t)kK;
#pragma unroll 1
                for (int t = 0; t < tiles; ++t) {
                    mbar_wait(&d_empty[rd.idx], rd.phase ^ 1u);
                    mbar_wait(&p_full[rp.idx], rp.phase);
                    tc_fence_after();
                    const uint64_t bdesc = smem_desc(smem_u32(smem + kOffPts + rp.idx * kPtBytes));
                    const uint32_t d_r = tmem_base + (uint32_t)(kColD0 + rd.idx * kColDStride);
                    const uint32_t d_j = d_r + (uint32_t)kPts;
                    DRB_UNROLL
                    for (int k = 0; k < ((DRB_TC_ABLATE & 1) ? 0 : kKSteps); ++k) {
                        // a K step covers 8 columns of the model operand: 8 TF32 words, or 16 BF16 elements packed
                        // two per column with the even K index in the low half (the packing is this file's one
                        // assumption that neither CUTLASS's tmem_frg layout algebra nor the host model pins down;
                        // the TF32 variant does not depend on it)
                        if (BF16) {
                            mma_bf16_ts(d_r, a_r + 8u * k, smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                            mma_bf16_ts(d_j, a_j + 8u * k, smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                        } else {
                            mma_tf32_ts(d_r, a_r + 8u * k, smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                            mma_tf32_ts(d_j, a_j + 8u * k, smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(&p_empty[rp.idx]);
                    mma_commit(&d_full[rd.idx]);
                    rp.advance(kStages);
                    rd.advance(2);
                }
                mma_commit(&a_empty[ra.idx]);
                ra.advance(2);
            }
        }
    } else if (warp < kWarpEpi0) {
        // ===== builders: thread = model (its TMEM lane); coefficient words straight into tensor memory =====
        const int quarter = warp & 3;
        const int i = quarter * 32 + lane;
        Ring ra;
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            const int mi = mt * kModels + i;
            float m[9];
            DRB_UNROLL
            for (int q = 0; q < 9; ++q) m[q] = mi < cnt ? __ldg(models + ((size_t)b * M + mi) * 9 + q) : 0.f;
            float cr[kFeat], cj[kFeat];
            uint32_t wr[kK], wj[kK];
            model_rows(m, mi < cnt, false, cr, cj);
            operand_row_words(cr, false, BF16, wr);
            operand_row_words(cj, false, BF16, wj);
            mbar_wait(&a_empty[ra.idx], ra.phase ^ 1u);
            __syncwarp();
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kColA0 + ra.idx * kColAStride);
            DRB_UNROLL
            for (int c = 0; c < kK / 16; ++c) {
                tmem_st16(ta + (uint32_t)(16 * c), wr + 16 * c);
                tmem_st16(ta + (uint32_t)(kK + 16 * c), wj + 16 * c);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ra.idx]);
            ra.advance(2);
        }
    } else {
        // ===== epilogue: lane = model, columns = correspondences =====
        constexpr int kParts = EPI / 4, kCols = kPts / kParts;     // 40 columns per warp (EPI = 8) or 20 (16)
        static_assert(kCols == 40 || kCols == 20, "column split");
        const int et = threadIdx.x - kWarpEpi0 * 32;
        const int quarter = warp & 3;
        const int cpart = (warp - kWarpEpi0) >> 2;
        Ring rd;
        int parity = 0;
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, parity ^= 1) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            const float th = 1.5f * __ldg(thr + b);
            const float nci = -1.f / (th * th);
            pk2 acc[4];
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) acc[i] = pk2_splat(0.f);
#pragma unroll 1
            for (int t = 0; t < tiles; ++t) {
                mbar_wait(&d_full[rd.idx], rd.phase);
                __syncwarp();
                tc_fence_after();
                const uint32_t tr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                    (uint32_t)(kColD0 + rd.idx * kColDStride + cpart * kCols);
                const uint32_t tj = tr + (uint32_t)kPts;
                uint32_t vr[kCols], vj[kCols];
                if constexpr (kCols == 40) {
                    tmem_ld32(tr, vr);
                    tmem_ld8(tr + 32u, vr + 32);
                    tmem_ld32(tj, vj);
                    tmem_ld8(tj + 32u, vj + 32);
                } else {
                    tmem_ld16(tr, vr);
                    tmem_ld4(tr + 16u, vr + 16);
                    tmem_ld16(tj, vj);
                    tmem_ld4(tj + 16u, vj + 16);
                }
                tmem_ld_wait();
                // the accumulator is in registers: hand it back before the arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[rd.idx]);
                rd.advance(2);
                // PAIR: one reciprocal per two neighbouring correspondences, (1/j0, 1/j1) = rcp(j0 j1) (j1, j0).  A row
                // past N has r = j = 0 and would take its neighbour with it (0 * inf), so when N is odd the last
                // tile -- the only place where a real and an absent correspondence share a pair -- takes the plain path.
                const bool paired = PAIR && !((N & 1) && t == tiles - 1);
                if (DRB_TC_ABLATE & 2) {
                    acc[0] = pk2_add(acc[0], pk2_make(__uint_as_float(vr[0] ^ vr[kCols - 1]), __uint_as_float(vj[0] ^ vj[kCols - 1])));
                } else if (paired) {
                    DRB_UNROLL
                    for (int i = 0; i < kCols / 2; ++i) {
                        const float j0 = __uint_as_float(vj[2 * i]), j1 = __uint_as_float(vj[2 * i + 1]);
                        float q0, q1;
                        const pk2 R = pk2_make(__uint_as_float(vr[2 * i]), __uint_as_float(vr[2 * i + 1]));
                        pk2_split(pk2_mul(R, R), q0, q1);
                        const float tn = rcp_approx(j0 * j1) * nci;
                        acc[i & 3] = pk2_add(acc[i & 3], pk2_make(fma_sat(q0 * j1, tn, 1.f), fma_sat(q1 * j0, tn, 1.f)));
                    }
                } else {
                    DRB_UNROLL
                    for (int i = 0; i < kCols / 2; ++i) {
                        const pk2 R = pk2_make(__uint_as_float(vr[2 * i]), __uint_as_float(vr[2 * i + 1]));
                        const pk2 IJ = pk2_make(rcp_approx(__uint_as_float(vj[2 * i])), rcp_approx(__uint_as_float(vj[2 * i + 1])));
                        float u0, u1;
                        pk2_split(pk2_mul(pk2_mul(R, R), IJ), u0, u1);
                        acc[i & 3] = pk2_add(acc[i & 3], pk2_make(fma_sat(u0, nci, 1.f), fma_sat(u1, nci, 1.f)));
                    }
                }
            }
            float lo, hi;
            pk2_split(pk2_add(pk2_add(acc[0], acc[1]), pk2_add(acc[2], acc[3])), lo, hi);
            float* pp = part + (size_t)parity * (4 * kModels);
            pp[cpart * kModels + quarter * 32 + lane] = lo + hi;
            asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
            if (et < kModels) {
                float score = pp[et];
                DRB_UNROLL
                for (int k = 1; k < kParts; ++k) score += pp[k * kModels + et];
                const int mi = mt * kModels + et;
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

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

static int sm_count() { return sm_count_current_device(); }

size_t workspace_bytes(int B, int N) { return (size_t)B * ((N + kPts - 1) / kPts) * kPtBytes; }

template <bool BF16, int EPI, bool PAIR>
int launch(const float* matches, const float* models, const int32_t* count, const int32_t* ids, const float* thr, int B,
           int M, int N, float* scores, unsigned long long* best_packed, uint32_t* images, cudaStream_t s) {
    static std::atomic<unsigned long long> configured{0};
    if (!ensure_dynamic_smem(score_msac_tc2_kernel<BF16, EPI, PAIR>, kSmemBytes, configured)) return DRB_ERR_CUDA;
    const int tiles = (N + kPts - 1) / kPts;
    msac_tc2_features_kernel<BF16><<<dim3(tiles, B), 96, 0, s>>>(matches, N, tiles, images);
    const long long max_units = (long long)B * ((M + kModels - 1) / kModels);
    const int grid = (int)(max_units < sm_count() ? max_units : sm_count());
    score_msac_tc2_kernel<BF16, EPI, PAIR><<<grid, threads_of(EPI), kSmemBytes, s>>>(images, models, count, ids, thr, B, M, N,
                                                                              tiles, scores, best_packed);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

// called by drb_score_msac_tc (score_tc.cu) for words + 64
int dispatch(bool bf16, bool e16, bool pair, const float* matches, const float* models, const int32_t* count,
             const int32_t* ids, const float* thr, int B, int M, int N, float* scores, unsigned long long* best_packed,
             uint32_t* images, cudaStream_t s) {
#define DRB_TC2_ARGS matches, models, count, ids, thr, B, M, N, scores, best_packed, images, s
#define DRB_TC2_PICK(BF, E) (pair ? launch<BF, E, true>(DRB_TC2_ARGS) : launch<BF, E, false>(DRB_TC2_ARGS))
    if (bf16) return e16 ? DRB_TC2_PICK(true, 16) : DRB_TC2_PICK(true, 8);
    return e16 ? DRB_TC2_PICK(false, 16) : DRB_TC2_PICK(false, 8);
#undef DRB_TC2_PICK
#undef DRB_TC2_ARGS
}

}  // namespace tc2
}  // namespace drb
