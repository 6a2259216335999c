// Soft-MSAC scoring on the 5th-generation tensor cores (ops.score_msac(kernel="tc_tf32" | "tc_bf16"); the
// default scorer of the pipelined service, engine.SERVICE_SCORER).
//
// Replaces the same reference lines as score.cu / score_stream.cu -- scorings/msac_score.py:12-55 (Sampson
// residuals of all N correspondences against all M models, soft inlier score) and the arg-max of
// ransac.py:114 -- with the same C-ABI contract (drb_score_msac_tc in include/drb.h).
//
// Why: the FP32 kernels spend ~20 FMA-pipe instructions per (model, correspondence) and sit at ~62 % of the
// FP32 peak.  Twelve of those instructions evaluate two polynomials in the
// correspondence's coordinates whose coefficients depend only on the model -- r = x2' M x1 and the Sampson
// denominator j -- i.e. a contraction over 15 monomials (msac_tc_layout.cuh).  Here that contraction runs on
// tcgen05 (operands split into TF32 or BF16 words, fp32 accumulation in tensor memory) and the CUDA cores keep
// only the epilogue r^2 / j -> clamp -> sum: 3 packed + 2 scalar FMA-pipe instructions and 2 reciprocals per
// two pairs.
//
// Two launches:
//   msac_tc_features_kernel   one thread per correspondence: the 48-word operand row (monomials split into
//                             words), written straight in the shared-memory image of its 128-row tile, so the
//                             scorer fetches a tile with ONE 24 KB bulk copy (no tensor map)
//   score_msac_tc_kernel      persistent, one CTA per SM, 12 warps:
//       warp 0      producer: cp.async.bulk of the correspondence tiles into a 4-stage ring
//       warp 1      allocates the 512 TMEM columns; one lane issues 6 x tcgen05.mma (128 x 256 x 8 tf32 or x 16
//                   bf16) per tile into one of two 256-column accumulators and commits to the mbarriers
//       warps 2-3   build the model operand (coefficient rows split into words) of the NEXT unit in the second
//                   B buffer while the current unit is being scored
//       warps 4-11  (4-19 in the 16-warp variant) epilogue: tcgen05.ld 32 lanes x 32 columns, r^2 * rcp(j), FFMA.SAT, per-thread sums over
//                   the unit's tiles, then a butterfly reduce-scatter over the 32 lanes and a fixed-order sum
//                   of the four lane quarters (the scores do not depend on the schedule)
//   A unit = (pair, 128 consecutive models); units are dealt round-robin to the CTAs.
//
// Measured on B200 (cfg2, 139 000 models x 2000 correspondences): 0.1075 ms for the pair-reciprocal BF16 variant
// (0.122 ms with one reciprocal per pair) against 0.215 ms for the FP32 work queue.  What bounds it (round-2
// ablations, DESIGN.md section 10, profiles/r2_tc_ablate.jsonl): the tensor side alone takes 0.085 ms -- 78 % of the
// dense BF16 rate this pool sustains under its power cap -- and the epilogue alone 0.095-0.10 ms (issue-bound: 305
// instructions per warp and tile); neither the XU nor the FMA pipe is the limiter.  Every variant is pinned to the
// fp64 oracle on the B200 (tests/test_gpu_score_tc.py); operand images / column mapping / descriptors are also
// checked on the host (tests/test_host_math.py::test_msac_tc_*).
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "device_cfg.cuh"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "msac_tc_layout.cuh"
#include "sampson.cuh"
#include "tc_ptx.cuh"
#include "tile_pipe.cuh"

// Ablation switches for profiles/microbench/tc_ablate.py (bit 0: no MMAs are issued, bit 1: the epilogue skips its
// arithmetic, bit 2: the epilogue skips tcgen05.ld, bit 3: a multiply stands in for MUFU.RCP, bit 4: the accumulator
// is handed back right after the tile's last tcgen05.ld instead of after its arithmetic, bit 5 / bit 6: four / three of
// the six K steps are issued -- what fewer tensor flops would buy, bit 7: the idle roles poll their barriers without
// the nanosleep back-off).  The product build is 0:
// every switch is a compile-time constant and the kernel's SASS does not change.
#ifndef DRB_TC_ABLATE
#define DRB_TC_ABLATE 0
#endif

// DRB_TC_HALF: the accumulator is handed over in HALVES of 128 columns (64 models) -- per tile two groups of six
// N = 128 MMAs, each with its own full / empty barrier pair and its own group of epilogue warps -- instead of as one
// 256-column unit: four half-buffers in flight instead of two buffers, so the MMAs of one half overlap the epilogue of
// the other and a slow warp holds back 128 columns, not 256.
#ifndef DRB_TC_HALF
#define DRB_TC_HALF 0
#endif

namespace drb {
namespace tc {

// Waits of the single-thread roles (producer, MMA issuer) and of the builders: poll, then sleep 100 ns.  mbar_wait
// re-issues mbarrier.try_wait as fast as the warp can -- a quarter of the kernel's executed instructions were such
// polls, issued on the sub-partitions the epilogue warps need; with the back-off the kernel is 2-3 % faster
// (0.1117 -> 0.1086 ms at cfg2; a suspend-time hint on try_wait: 0.1097).  The epilogue warps keep the plain wait:
// their wake-up latency is on the critical path.  DRB_TC_ABLATE bit 7 restores the plain wait everywhere.
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    if (DRB_TC_ABLATE & 128) {
        mbar_wait(bar, parity);
        return;
    }
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
        __nanosleep(100);
    }
}

constexpr int kWarpProducer = 0;
constexpr int kWarpMma = 1;
constexpr int kWarpBuild0 = 2;       // warps 2, 3
constexpr int kBuildThreads = 64;
constexpr int kWarpEpi0 = 4;         // epilogue warps 4 .. 4 + EPI - 1 (EPI = 8 or 16)
constexpr int threads_of(int epi_warps) { return (kWarpEpi0 + epi_warps) * 32; }
constexpr int kTmemCols = 512;       // two accumulators of kTileN columns
constexpr int kMaxPairs = 1024;

// dynamic shared memory carve-up (bytes) for STAGES correspondence stages
template <int STAGES>
struct Carve {
    static constexpr int kOffA = 0;
    static constexpr int kOffB = kOffA + STAGES * kABytes;            //  98304 with four stages
    static constexpr int kOffBars = kOffB + 2 * kBBytes;              // 196608
    static constexpr int kNumBars = 2 * STAGES + 4 + 4 + 2 + 2;       // a_full/a_empty, d_full, d_empty (per half), b_full, b_empty
    static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
    static constexpr int kOffPrefix = kOffTmemPtr + 16;
    static constexpr int kOffPart = kOffPrefix + (kMaxPairs + 1) * 4 + 12;
    static constexpr int kSmemBytes = kOffPart + 2 * 4 * kTileModels * 4;   // part[unit parity][lane quarter][model of the tile]
    static_assert(kOffPart % 16 == 0, "alignment");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};
// The "slim" build (words + 256): 128 registers per thread instead of 168 and three correspondence stages instead of
// four, so that a CTA leaves 16 K registers and ~50 KB of shared memory of its SM free -- room for one five-point CTA
// (160 threads x 96 registers, 43 KB) of the NEXT batch on another stream while this one is being scored.
constexpr int stages_of(bool slim) { return slim ? 3 : 4; }

// ---- launch 1: correspondences -> operand images ------------------------------------------------------
// images[b][t] = the kABytes image of correspondences [128 t, 128 t + 128) of pair b; rows past N are zero
// (the scorer's epilogue masks them).
// PADFLAG (the folded variant, msac_tc_layout.cuh): a row past N -- or a correspondence that is not finite, which the
// other variants turn into 0 through FFMA.SAT -- is (0, ..., 0, 1): the sixteenth slot makes every model answer
// (r, j') = (1e18, -1), a term of exactly 0.
template <bool BF16, bool PADFLAG>
__global__ void __launch_bounds__(kTileM)
msac_tc_features_kernel(const float* __restrict__ matches, int N, int tiles, uint32_t* __restrict__ images) {
    const int b = blockIdx.y, t = blockIdx.x, row = threadIdx.x;
    const int n = t * kTileM + row;
    uint32_t row48[kK];
    bool real = n < N;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (real) {
        p = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)b * N + n);
        if (PADFLAG) real = fabsf(p.x) <= 1.0e18f && fabsf(p.y) <= 1.0e18f && fabsf(p.z) <= 1.0e18f && fabsf(p.w) <= 1.0e18f;
    }
    if (real) {
        float f[kFeat];
        features(p.x, p.y, p.z, p.w, f);
        operand_row_words(f, true, BF16, row48);
    } else if (PADFLAG) {
        float f[kFeat];
        DRB_UNROLL
        for (int k = 0; k < kFeat; ++k) f[k] = 0.f;
        operand_row_words(f, true, BF16, row48, 1.f);
    } else {
        DRB_UNROLL
        for (int k = 0; k < kK; ++k) row48[k] = 0u;
    }
    uint32_t* img = images + ((size_t)b * tiles + t) * (kABytes / 4);
    DRB_UNROLL
    for (int c = 0; c < kK / 4; ++c)
        *reinterpret_cast<uint4*>(img + image_index(row, 4 * c)) =
            make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
}

// ---- launch 2 -----------------------------------------------------------------------------------------
// PAIR: 0 one reciprocal per pair; 1 one per two neighbouring models; 2 the folded form of 1 (msac_tc_layout.cuh)
template <bool BF16, int PAIR, int EPI, bool SLIM>
__global__ void __launch_bounds__(SLIM ? 512 : threads_of(EPI), 1)
score_msac_tc_kernel(const uint32_t* __restrict__ images, const float* __restrict__ models,
                     const int32_t* __restrict__ count, const int32_t* __restrict__ ids, const float* __restrict__ thr,
                     int B, int M, int N, int tiles, float* __restrict__ scores,
                     unsigned long long* __restrict__ best_packed) {
    constexpr int kStagesA = stages_of(SLIM);
    using C = Carve<kStagesA>;
    constexpr int kOffA = C::kOffA, kOffB = C::kOffB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + kStagesA;
    uint64_t* d_full = bars + 2 * kStagesA;      // [buffer][half]: index 2 * buffer + half (half 0 alone unless DRB_TC_HALF)
    uint64_t* d_empty = d_full + 4;
    uint64_t* b_full = d_empty + 4;
    uint64_t* b_empty = b_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::kOffTmemPtr);
    int* prefix = reinterpret_cast<int*>(smem + C::kOffPrefix);
    float* part = reinterpret_cast<float*>(smem + C::kOffPart);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ---------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStagesA; ++i) {
            mbar_init(&a_full[i], 1);     // the producer's arrive.expect_tx
            mbar_init(&a_empty[i], 1);    // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            for (int h = 0; h < 2; ++h) {
                mbar_init(&d_full[2 * i + h], 1);                             // tcgen05.commit
                mbar_init(&d_empty[2 * i + h], DRB_TC_HALF ? EPI / 2 : EPI);  // one arrival per epilogue warp (of the half)
            }
            mbar_init(&b_full[i], kBuildThreads / 32); // one arrival per builder warp
            mbar_init(&b_empty[i], 1);                 // tcgen05.commit
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWarpEpi0) unit_prefix(prefix, count, B, M, kTileModels, lane);
    if (warp == kWarpMma) tmem_alloc(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_units = prefix[B];

    if (warp == kWarpProducer) {
        // ===== producer: correspondence tiles of every unit of this CTA, in unit order =====
        if (lane == 0) {
            Ring ra;
#pragma unroll 1
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int b, mt;
                unit_of(prefix, B, u, b, mt);
                const uint32_t* src = images + (size_t)b * tiles * (kABytes / 4);
#pragma unroll 1
                for (int t = 0; t < tiles; ++t) {
                    mbar_wait_idle(&a_empty[ra.idx], ra.phase ^ 1u);
                    mbar_expect_tx(&a_full[ra.idx], kABytes);
                    bulk_g2s(smem + kOffA + ra.idx * kABytes, src + (size_t)t * (kABytes / 4), kABytes, &a_full[ra.idx]);
                    ra.advance(kStagesA);
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = BF16 ? instr_desc_bf16() : instr_desc();
            Ring ra, rd, rb;
#pragma unroll 1
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                mbar_wait_idle(&b_full[rb.idx], rb.phase);
                const uint64_t bdesc = smem_desc(smem_u32(smem + kOffB + rb.idx * kBBytes));
#pragma unroll 1
                for (int t = 0; t < tiles; ++t) {
                    const uint64_t adesc = smem_desc(smem_u32(smem + kOffA + ra.idx * kABytes));
                    if (DRB_TC_HALF) {
                        // idesc with N = 128: the N field (bits 17-22, N >> 3) of the 256-column descriptor halved
                        const uint32_t idesc_h = (idesc & ~(0x3fu << 17)) | ((uint32_t)(kTileN / 2 >> 3) << 17);
                        DRB_UNROLL
                        for (int h = 0; h < 2; ++h) {
                            mbar_wait_idle(&d_empty[2 * rd.idx + h], rd.phase ^ 1u);
                            if (h == 0) mbar_wait_idle(&a_full[ra.idx], ra.phase);
                            tc_fence_after();
                            const uint32_t d = tmem_base + (uint32_t)(rd.idx * kTileN + h * (kTileN / 2));
                            const uint64_t bd = bdesc + (uint64_t)((h * (kBBytes / 2)) >> 4);   // rows 128 h .. of the B image
                            DRB_UNROLL
                            for (int k = 0; k < kKSteps; ++k) {
                                if (BF16) mma_bf16(d, smem_desc_kstep(adesc, k), smem_desc_kstep(bd, k), idesc_h, k > 0 ? 1u : 0u);
                                else mma_tf32(d, smem_desc_kstep(adesc, k), smem_desc_kstep(bd, k), idesc_h, k > 0 ? 1u : 0u);
                            }
                            if (h == 1) mma_commit(&a_empty[ra.idx]);
                            mma_commit(&d_full[2 * rd.idx + h]);
                        }
                    } else {
                    mbar_wait_idle(&d_empty[2 * rd.idx], rd.phase ^ 1u);
                    mbar_wait_idle(&a_full[ra.idx], ra.phase);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(rd.idx * kTileN);
                    DRB_UNROLL
                    for (int k = 0; k < ((DRB_TC_ABLATE & 1) ? 0 : (DRB_TC_ABLATE & 32) ? 4 : (DRB_TC_ABLATE & 64) ? 3 : kKSteps); ++k) {
                        if (BF16) mma_bf16(d, smem_desc_kstep(adesc, k), smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                        else mma_tf32(d, smem_desc_kstep(adesc, k), smem_desc_kstep(bdesc, k), idesc, k > 0 ? 1u : 0u);
                    }
                    mma_commit(&a_empty[ra.idx]);   // the stage may be refilled once these MMAs have read it
                    mma_commit(&d_full[2 * rd.idx]);    // the accumulator is complete
                    }
                    ra.advance(kStagesA);
                    rd.advance(2);
                }
                mma_commit(&b_empty[rb.idx]);       // the model operand may be rebuilt
                rb.advance(2);
            }
        }
    } else if (warp < kWarpEpi0) {
        // ===== builders: the model operand of every unit, one buffer ahead of the MMAs =====
        const int bt = threadIdx.x - kWarpBuild0 * 32;   // 0 .. 63
        Ring rb;
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            mbar_wait_idle(&b_empty[rb.idx], rb.phase ^ 1u);
            uint32_t* img = reinterpret_cast<uint32_t*>(smem + kOffB + rb.idx * kBBytes);
            DRB_UNROLL
            for (int rep = 0; rep < kTileModels / kBuildThreads; ++rep) {
                const int i = bt + rep * kBuildThreads;   // model of the tile
                const int mi = mt * kTileModels + i;
                float m[9];
                DRB_UNROLL
                for (int q = 0; q < 9; ++q) m[q] = mi < cnt ? __ldg(models + ((size_t)b * M + mi) * 9 + q) : 0.f;
                float cr[kFeat], cj[kFeat], cr15 = 0.f, cj15 = 0.f;
                uint32_t row48[kK];
                if (PAIR == 2) {
                    const float th = 1.5f * __ldg(thr + b);
                    model_rows_folded(m, mi < cnt, -(th * th), cr, cj, cr15, cj15);
                } else {
                    model_rows(m, mi < cnt, PAIR != 0, cr, cj);
                }
                operand_row_words(cr, false, BF16, row48, cr15);
                DRB_UNROLL
                for (int c = 0; c < kK / 4; ++c)
                    *reinterpret_cast<uint4*>(img + image_index(column_r(i), 4 * c)) =
                        make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
                operand_row_words(cj, false, BF16, row48, cj15);
                DRB_UNROLL
                for (int c = 0; c < kK / 4; ++c)
                    *reinterpret_cast<uint4*>(img + image_index(PAIR ? column_j_swapped(i) : column_j(i), 4 * c)) =
                        make_uint4(row48[4 * c], row48[4 * c + 1], row48[4 * c + 2], row48[4 * c + 3]);
            }
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[rb.idx]);
            rb.advance(2);
        }
    } else {
        // ===== epilogue =====
        // EPI / 4 warps share a lane quarter; each takes kCols = 256 / (EPI / 4) consecutive columns of the
        // accumulator = kCols / 2 models, in kChunks loads of 32 columns (8 model pairs each)
        constexpr int kParts = EPI / 4, kCols = kTileN / kParts, kChunks = kCols / 32, kAcc = kChunks * 8;
        const int et = threadIdx.x - kWarpEpi0 * 32;   // 0 .. 32 EPI - 1
        const int quarter = warp & 3;                  // the TMEM lanes this warp may read: 32 quarter .. + 31
        const int half = (warp - kWarpEpi0) >> 2;      // this warp's column part: columns kCols half .. + kCols - 1
        const int dh = DRB_TC_HALF ? (half * kCols) / (kTileN / 2) : 0;   // which half of the accumulator that is
        Ring rd;
        int parity = 0;
#pragma unroll 1
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, parity ^= 1) {
            int b, mt;
            unit_of(prefix, B, u, b, mt);
            const int cnt = count ? min(__ldg(count + b), M) : M;
            const float th = 1.5f * __ldg(thr + b);
            const float nci = -1.f / (th * th);
            pk2 acc[kAcc];
            DRB_UNROLL
            for (int i = 0; i < kAcc; ++i) acc[i] = pk2_splat(0.f);
#pragma unroll 1
            for (int t = 0; t < tiles; ++t) {
                mbar_wait(&d_full[2 * rd.idx + dh], rd.phase);
                __syncwarp();          // tcgen05.ld is .sync.aligned: the warp must be converged
                tc_fence_after();
                // a row past N contributes 0: max(0, min(1, u * nci + 0)) with u * nci <= 0 (or NaN -> 0)
                const float one = (PAIR == 2 || t * kTileM + quarter * 32 + lane < N) ? 1.f : 0.f;   // (unused when PAIR == 2)
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(rd.idx * kTileN + half * kCols);
                DRB_UNROLL
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t v[32];
                    if (DRB_TC_ABLATE & 4) {
                        DRB_UNROLL
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(1.f + (float)(t + i + lane));
                    } else {
                        tmem_ld32(taddr + (uint32_t)(c * 32), v);
                        tmem_ld_wait();
                    }
                    if (c == kChunks - 1 && (DRB_TC_ABLATE & 16)) {
                        // measured and not kept: handing the accumulator back here, before the arithmetic of the last
                        // chunk, is 2 % slower (0.1096 vs 0.1075 ms, profiles/r2_tc_ablate.jsonl)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_empty[2 * rd.idx + dh]);
                    }
                    if (DRB_TC_ABLATE & 2) {
                        acc[c * 8] = pk2_add(acc[c * 8], pk2_make(__uint_as_float(v[0]), __uint_as_float(v[31])));
                        continue;
                    }
                    DRB_UNROLL
                    for (int q = 0; q < 8; ++q) {
                        const pk2 R = pk2_make(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
                        const pk2 R2 = pk2_mul(R, R);
                        if (PAIR == 2) {
                            // columns (r0, r1, j1', j0'), j' = -T j: term = 1 + t max(r^2 j_other', -p), t = 1 / p, p = j0' j1'
                            const float ja = __uint_as_float(v[4 * q + 2]), jb = __uint_as_float(v[4 * q + 3]);
                            const float pp = ja * jb;
                            const float tt = (DRB_TC_ABLATE & 8) ? pp * 0.37f : rcp_approx(pp);
                            float w0, w1, a0, a1;
                            pk2_split(pk2_mul(R2, pk2_make(ja, jb)), w0, w1);
                            pk2_split(acc[c * 8 + q], a0, a1);
                            acc[c * 8 + q] = pk2_make(fmaf(tt, fmaxf(w0, -pp), a0), fmaf(tt, fmaxf(w1, -pp), a1));
                        } else if (PAIR) {
                            // columns (r0, r1, j1, j0): (1/j0, 1/j1) = rcp(j0 j1) (j1, j0) -- one reciprocal per two pairs
                            const float ja = __uint_as_float(v[4 * q + 2]), jb = __uint_as_float(v[4 * q + 3]);
                            const float tn = ((DRB_TC_ABLATE & 8) ? (ja * jb) * 0.37f : rcp_approx(ja * jb)) * nci;
                            float w0, w1;
                            pk2_split(pk2_mul(R2, pk2_make(ja, jb)), w0, w1);
                            acc[c * 8 + q] = pk2_add(acc[c * 8 + q], pk2_make(fma_sat(w0, tn, one), fma_sat(w1, tn, one)));
                        } else {
                            const pk2 IJ = pk2_make(rcp_approx(__uint_as_float(v[4 * q + 2])), rcp_approx(__uint_as_float(v[4 * q + 3])));
                            float u0, u1;
                            pk2_split(pk2_mul(R2, IJ), u0, u1);
                            acc[c * 8 + q] = pk2_add(acc[c * 8 + q], pk2_make(fma_sat(u0, nci, one), fma_sat(u1, nci, one)));
                        }
                    }
                }
                if (!(DRB_TC_ABLATE & 16)) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[2 * rd.idx + dh]);
                }
                rd.advance(2);
            }
            // ---- sum over the 128 lanes: butterfly reduce-scatter inside the warp, then the four quarters -----
            float a[2 * kAcc];
            DRB_UNROLL
            for (int i = 0; i < kAcc; ++i) pk2_split(acc[i], a[2 * i], a[2 * i + 1]);
            if (PAIR == 2) {
                // the "1 +" of the row this thread met in every tile (a flagged row answered exactly -1)
                const float rows = (float)tiles;
                DRB_UNROLL
                for (int i = 0; i < 2 * kAcc; ++i) a[i] += rows;
            }
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
            // this lane now owns kPer = kAcc / 16 consecutive models of its part: kPer lane .. kPer lane + kPer - 1
            // (lane quarter `quarter`); part[parity][quarter][model of the tile]
            constexpr int kPer = kAcc / 16;
            float* pp = part + (size_t)parity * (4 * kTileModels);
            DRB_UNROLL
            for (int i = 0; i < kPer; ++i) pp[quarter * kTileModels + half * (kCols / 2) + kPer * lane + i] = a[i];
            asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
            if (et < kTileModels) {
                float score = ((pp[0 * kTileModels + et] + pp[1 * kTileModels + et]) + pp[2 * kTileModels + et]) +
                              pp[3 * kTileModels + et];
                if (PAIR == 2) score = fmaxf(score, 0.f);    // -inf / NaN of a degenerate denominator (p underflows): 0
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
            // `part` alternates between two halves by unit parity: the barrier of the next unit orders these
            // reads before the writes of the unit after it
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

static int tc_sm_count() { return sm_count_current_device(); }

// the operand images with `tiles` (>= ceil(N / 128)) tiles per pair -- rows past N are zero -- for score_tc_pair.cu,
// whose SM pairs consume the tiles two at a time
int launch_features(bool bf16, const float* matches, int B, int N, int tiles, uint32_t* images, cudaStream_t s) {
    if (bf16) msac_tc_features_kernel<true, false><<<dim3(tiles, B), kTileM, 0, s>>>(matches, N, tiles, images);
    else msac_tc_features_kernel<false, false><<<dim3(tiles, B), kTileM, 0, s>>>(matches, N, tiles, images);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}

}  // namespace tc
}  // namespace drb

using namespace drb;

namespace drb {
namespace tcp {   // score_tc_pair.cu: two SMs per tile (cta_group::2)
int dispatch(bool bf16, const float* matches, const float* models, const int32_t* count, const int32_t* ids,
             const float* thr, int B, int M, int N, float* scores, unsigned long long* best_packed, uint32_t* images,
             cudaStream_t s);
}
namespace tc2 {   // score_tc2.cu: the model-stationary arrangement
size_t workspace_bytes(int B, int N);
int dispatch(bool bf16, bool e16, bool pair, const float* matches, const float* models, const int32_t* count,
             const int32_t* ids, const float* thr, int B, int M, int N, float* scores, unsigned long long* best_packed,
             uint32_t* images, cudaStream_t s);
}  // namespace tc2
}  // namespace drb

extern "C" size_t drb_score_msac_tc_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    const size_t v1 = (size_t)B * (2 * ((N + 2 * tc::kTileM - 1) / (2 * tc::kTileM))) * tc::kABytes;   // an even tile count (+ 512)
    const size_t v2 = tc2::workspace_bytes(B, N);     // tiles of 80 correspondences: a different padding
    return v1 > v2 ? v1 : v2;
}

namespace drb {
namespace tc {
template <bool BF16, int PAIR, int EPI, bool SLIM>
static int launch(const float* matches, const float* models, const int32_t* count, const int32_t* ids, const float* thr,
                  int B, int M, int N, float* scores, unsigned long long* best_packed, uint32_t* images, cudaStream_t s) {
    static std::atomic<unsigned long long> configured{0};
    constexpr int kSmemBytes = Carve<stages_of(SLIM)>::kSmemBytes;
    if (!ensure_dynamic_smem(score_msac_tc_kernel<BF16, PAIR, EPI, SLIM>, kSmemBytes, configured)) return DRB_ERR_CUDA;
    const int tiles = (N + kTileM - 1) / kTileM;
    msac_tc_features_kernel<BF16, PAIR == 2><<<dim3(tiles, B), kTileM, 0, s>>>(matches, N, tiles, images);
    const long long max_units = (long long)B * ((M + kTileModels - 1) / kTileModels);
    const int grid = (int)(max_units < tc_sm_count() ? max_units : tc_sm_count());
    score_msac_tc_kernel<BF16, PAIR, EPI, SLIM><<<grid, threads_of(EPI), kSmemBytes, s>>>(images, models, count, ids, thr, B, M, N, tiles, scores,
                                                                  best_packed);
    return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA;
}
}  // namespace tc
}  // namespace drb

extern "C" int drb_score_msac_tc(const float* matches, const float* models, const int32_t* count, const int32_t* ids,
                                 const float* thr, int B, int M, int N, int words, float* scores,
                                 unsigned long long* best_packed, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!matches || !models || !thr || !best_packed || !workspace) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || B > tc::kMaxPairs || M <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    if (workspace_bytes < drb_score_msac_tc_workspace_bytes(B, N) || (reinterpret_cast<uintptr_t>(workspace) & 127) ||
        (reinterpret_cast<uintptr_t>(matches) & 15))
        return DRB_ERR_BAD_SHAPE;
    // words: 2 = TF32 x 2, 3 = BF16 x 3; + 16 = one reciprocal per model pair; + 32 = 16 epilogue warps instead
    // of 8; + 64 = the model-stationary arrangement of score_tc2.cu; + 128 = folded threshold; + 256 = slim build
    // (all measured on the B200: DESIGN.md section 10)
    const int split = words & 15;
    const bool pair = (words & 16) != 0, e16 = (words & 32) != 0, v2 = (words & 64) != 0, fold = (words & 128) != 0,
               slim = (words & 256) != 0, two_sm = (words & 512) != 0;
    if ((split != 2 && split != 3) || (words & ~1023) || (fold && (!pair || v2)) || (slim && (!pair || fold || e16 || v2)) ||
        (two_sm && (!pair || fold || e16 || v2 || slim)))
        return DRB_ERR_UNSUPPORTED;
    uint32_t* images = reinterpret_cast<uint32_t*>(workspace);
    cudaStream_t s = (cudaStream_t)stream;
    if (two_sm) return tcp::dispatch(split == 3, matches, models, count, ids, thr, B, M, N, scores, best_packed,
                                     reinterpret_cast<uint32_t*>(workspace), (cudaStream_t)stream);
    if (v2) return tc2::dispatch(split == 3, e16, pair, matches, models, count, ids, thr, B, M, N, scores, best_packed, images, s);
#define DRB_TC_ARGS matches, models, count, ids, thr, B, M, N, scores, best_packed, images, s
#define DRB_TC_PICK(BF, PR) (e16 ? tc::launch<BF, PR, 16, false>(DRB_TC_ARGS) : tc::launch<BF, PR, 8, false>(DRB_TC_ARGS))
    if (slim) return split == 3 ? tc::launch<true, 1, 8, true>(DRB_TC_ARGS) : tc::launch<false, 1, 8, true>(DRB_TC_ARGS);
    if (fold) return split == 3 ? DRB_TC_PICK(true, 2) : DRB_TC_PICK(false, 2);
    if (pair) return split == 3 ? DRB_TC_PICK(true, 1) : DRB_TC_PICK(false, 1);
    return split == 3 ? DRB_TC_PICK(true, 0) : DRB_TC_PICK(false, 0);
#undef DRB_TC_PICK
#undef DRB_TC_ARGS
}
