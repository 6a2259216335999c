// Operand construction for the tensor-core soft-MSAC scorer (score_tc.cu): the residual numerator r and the
// Sampson denominator j of scorings/msac_score.py:26-42 written as ONE contraction over monomials of the
// correspondence, so that a 128 x 256 tile of (correspondence, model) pairs is a tcgen05 MMA:
//
//   r(n, m) = x2' M x1                         = <phi(n), cr(m)>
//   j(n, m) = (M x1)_0^2 + (M x1)_1^2 + (M' x2)_0^2 + (M' x2)_1^2 = <phi(n), cj(m)>
//   phi(n)  = (x2 x1, x2 y1, y2 x1, y2 y1, x1^2, x1 y1, y1^2, x2^2, x2 y2, y2^2, x1, y1, x2, y2, 1)
//
// The tensor core multiplies TF32 (10-bit mantissa) operands, which is far too coarse for r against a threshold
// of ~1e-3, so every operand is split into two TF32 words v = hi + lo and three of the four partial products
// are accumulated (3xTF32, error ~2^-22 relative -- measured 3-5e-5 relative on the scores against fp64, the
// fp32 formula gives 1-2e-5):  K = 48 = 3 blocks of 16 (15 monomials + one zero):
//
//   A (correspondences) = [ hi | lo | hi ],   B (models) = [ hi | hi | lo ]   ->   A B' = hi hi + lo hi + hi lo
//
// Everything here is DRB_HD so that tests/hostcheck can build the very same images on the host, decode them
// with the descriptor fields the kernel uses and check the scores against the oracle without a GPU
// (tests/test_host_math.py::test_msac_tc_*).
#pragma once

#include <stdint.h>
#include <string.h>

#include "drb_common.cuh"

namespace drb {
namespace tc {

constexpr int kFeat = 15;                      // monomials
constexpr int kBlk = 16;                       // one K block: the monomials + one zero
constexpr int kK = 3 * kBlk;                   // 48
constexpr int kMmaK = 8;                       // K of one tf32 tcgen05.mma (32 bytes)
constexpr int kKSteps = kK / kMmaK;            // 6
constexpr int kTileM = 128;                    // correspondences per tile = MMA M = TMEM lanes
constexpr int kTileModels = 128;               // models per tile
constexpr int kTileN = 2 * kTileModels;        // MMA N: an r column and a j column per model
// Canonical K-major, no-swizzle shared-memory image of an operand (cute::UMMA "INTERLEAVE" layout): 8 rows x
// 16 bytes form one 128-byte core matrix; the core matrices of one 8-row group follow each other along K
// (leading byte offset), the groups follow each other along M/N (stride byte offset).
constexpr int kLBO = 128;
constexpr int kSBO = (kK / 4) * kLBO;          // 1536
constexpr int kABytes = (kTileM / 8) * kSBO;   // 24576: one tile of correspondences
constexpr int kBBytes = (kTileN / 8) * kSBO;   // 49152: one tile of models

// float index of element (row, k) inside an operand image
DRB_HD int image_index(int row, int k) { return ((row >> 3) * kSBO + (k >> 2) * kLBO + (row & 7) * 16) / 4 + (k & 3); }

// D column (= B row) of the r / j value of model i of a tile: columns come in groups of four
// (r_2g, r_2g+1, j_2g, j_2g+1) so that the epilogue finds two models side by side in an aligned register pair
DRB_HD int column_r(int i) { return 4 * (i >> 1) + (i & 1); }
DRB_HD int column_j(int i) { return 4 * (i >> 1) + 2 + (i & 1); }
// Pair-reciprocal variant: the two j columns of a group are SWAPPED, (r_2g, r_2g+1, j_2g+1, j_2g), so that the
// epilogue's (1/j_2g, 1/j_2g+1) = rcp(j_2g j_2g+1) * (j_2g+1, j_2g) needs one reciprocal and no register moves
DRB_HD int column_j_swapped(int i) { return 4 * (i >> 1) + 2 + ((i & 1) ^ 1); }

// ---- descriptors ------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units,
// version 1 (Blackwell), no swizzle.
DRB_HD uint64_t smem_desc(uint32_t smem_byte_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_byte_addr >> 4) & 0x3fff);
    d |= (uint64_t)((kLBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((kSBO >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                    // version
    // base offset (bits 49-51) = 0, LBO mode (bit 52) = 0, layout type (bits 61-63) = 0: SWIZZLE_NONE
    return d;
}
// descriptor of the same operand advanced by `step` MMA K steps (two 16-byte chunks each)
DRB_HD uint64_t smem_desc_kstep(uint64_t desc, int step) { return desc + (uint64_t)((2 * kLBO * step) >> 4); }

// Instruction descriptor (cute::UMMA::InstrDescriptor) of kind::tf32, fp32 accumulate, K-major A and B
DRB_HD uint32_t instr_desc() {
    uint32_t d = 0;
    d |= 1u << 4;                              // D format: F32
    d |= 2u << 7;                              // A format: TF32
    d |= 2u << 10;                             // B format: TF32
    // bits 13/14: no negate; bits 15/16: A, B K-major
    d |= (uint32_t)(kTileN >> 3) << 17;
    d |= (uint32_t)(kTileM >> 4) << 24;
    return d;
}

// ---- TF32 split --------------------------------------------------------------------------------------
// Round to the nearest TF32 (ties away, what cvt.rna.tf32.f32 does); the result is an fp32 with 13 zero bits.
DRB_HD float tf32_round(float x) {
#if defined(__CUDA_ARCH__)
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;   // not Inf / NaN
    float r;
    memcpy(&r, &u, 4);
    return r;
#endif
}
DRB_HD void tf32_split(float x, float& hi, float& lo) {
    hi = tf32_round(x);
    lo = tf32_round(x - hi);   // x - hi is exact in fp32
}

// ---- monomials of a correspondence, coefficient rows of a model -----------------------------------------
DRB_HD void features(float x1, float y1, float x2, float y2, float* f) {
    f[0] = x2 * x1;
    f[1] = x2 * y1;
    f[2] = y2 * x1;
    f[3] = y2 * y1;
    f[4] = x1 * x1;
    f[5] = x1 * y1;
    f[6] = y1 * y1;
    f[7] = x2 * x2;
    f[8] = x2 * y2;
    f[9] = y2 * y2;
    f[10] = x1;
    f[11] = y1;
    f[12] = x2;
    f[13] = y2;
    f[14] = 1.f;
}
// m = 3x3 row-major, x2' M x1
DRB_HD void coefficients(const float* m, float* cr, float* cj) {
    cr[0] = m[0];
    cr[1] = m[1];
    cr[2] = m[3];
    cr[3] = m[4];
    cr[4] = cr[5] = cr[6] = cr[7] = cr[8] = cr[9] = 0.f;
    cr[10] = m[6];
    cr[11] = m[7];
    cr[12] = m[2];
    cr[13] = m[5];
    cr[14] = m[8];
    cj[0] = cj[1] = cj[2] = cj[3] = 0.f;
    // (m0 x1 + m1 y1 + m2)^2 + (m3 x1 + m4 y1 + m5)^2
    cj[4] = m[0] * m[0] + m[3] * m[3];
    cj[5] = 2.f * (m[0] * m[1] + m[3] * m[4]);
    cj[6] = m[1] * m[1] + m[4] * m[4];
    cj[10] = 2.f * (m[0] * m[2] + m[3] * m[5]);
    cj[11] = 2.f * (m[1] * m[2] + m[4] * m[5]);
    // (m0 x2 + m3 y2 + m6)^2 + (m1 x2 + m4 y2 + m7)^2
    cj[7] = m[0] * m[0] + m[1] * m[1];
    cj[8] = 2.f * (m[0] * m[3] + m[1] * m[4]);
    cj[9] = m[3] * m[3] + m[4] * m[4];
    cj[12] = 2.f * (m[0] * m[6] + m[1] * m[7]);
    cj[13] = 2.f * (m[3] * m[6] + m[4] * m[7]);
    cj[14] = (m[2] * m[2] + m[5] * m[5]) + (m[6] * m[6] + m[7] * m[7]);
}

// Coefficient rows of model i of a tile as the scorer's builder warps write them.  With the pair reciprocal a
// model's 1/j is computed from the product with its neighbour's j, so a column that is not a real model must not
// poison its neighbour: an absent model (past the pair's count) becomes r = 0, j = 1, and a model with a
// non-finite coefficient becomes r = 1e18, j = 1 (its own terms clamp to 0 exactly as NaN terms do, its
// neighbour's are untouched).  Without the pair reciprocal absent models are all-zero rows and NaN stays NaN.
DRB_HD void model_rows(const float* m, bool present, bool pair, float* cr, float* cj) {
    bool finite = true;
    DRB_UNROLL
    for (int q = 0; q < 9; ++q) finite = finite && (t_abs(m[q]) <= 3.0e38f);   // false for NaN and Inf
    if (present && (finite || !pair)) {
        coefficients(m, cr, cj);
        return;
    }
    DRB_UNROLL
    for (int i = 0; i < kFeat; ++i) cr[i] = cj[i] = 0.f;
    if (pair) {
        cj[kFeat - 1] = 1.f;
        if (present) cr[kFeat - 1] = 1.0e18f;
    }
}

// "Folded" pair variant (words + 128): the threshold moves into the denominator rows, j' = -(1.5 thr)^2 j < 0, so that
// the soft inlier term of a pair of neighbouring models becomes
//     max(0, 1 - r0^2 / (j0 T)) = 1 + t max(r0^2 j1', -p),   p = j0' j1' > 0,  t = 1 / p
// -- per two pairs FMUL2, FMUL, MUFU.RCP, FMUL2, 2 FMNMX (ALU pipe), 2 FFMA that accumulate: 7 FMA-pipe cycles
// instead of 10, no separate clamp and no separate add.  The sixteenth K slot (zero in the other variants) flags the
// rows that must not count (past N, or a correspondence that is not finite): their features are (0, ..., 0, 1) and
// every model carries (1e18, -1) there, so such a row gives r = 1e18, j' = -1, p = 1, t = 1, max(-1e36, -1) = -1:
// exactly -1, which cancels the "1 +" that the epilogue adds once per thread and tile at the end of a unit -- no
// mask in the inner loop.  Absent / non-finite / all-zero models become r = 0 | 1e18, j' = -1 (the sign matters: p
// must stay positive).
DRB_HD void model_rows_folded(const float* m, bool present, float jscale, float* cr, float* cj, float& cr15, float& cj15) {
    bool finite = true;
    DRB_UNROLL
    for (int q = 0; q < 9; ++q) finite = finite && (t_abs(m[q]) <= 3.0e38f);   // false for NaN and Inf
    cr15 = 1.0e18f;
    cj15 = -1.f;
    if (present && finite) {
        coefficients(m, cr, cj);
        float mass = 0.f;
        DRB_UNROLL
        for (int i = 0; i < kFeat; ++i) mass += t_abs(cj[i]);
        if (mass > 0.f) {                     // j is not identically zero
            DRB_UNROLL
            for (int i = 0; i < kFeat; ++i) cj[i] *= jscale;
            return;
        }
    }
    DRB_UNROLL
    for (int i = 0; i < kFeat; ++i) cr[i] = cj[i] = 0.f;
    cj[kFeat - 1] = -1.f;
    if (present) cr[kFeat - 1] = 1.0e18f;
}

// One operand row: the three K blocks of 16 (v = 15 values + v15, the sixteenth slot), A side [hi | lo | hi], B side
// [hi | hi | lo].  `row48` receives the 48 floats in K order.
DRB_HD void operand_row(const float* v, bool a_side, float* row48, float v15 = 0.f) {
    DRB_UNROLL
    for (int i = 0; i < kBlk; ++i) {
        float hi = 0.f, lo = 0.f;
        if (i < kFeat) tf32_split(v[i], hi, lo);
        else hi = tf32_round(v15);             // 0, 1, -1 or 1e18: one TF32 word is enough
        row48[i] = hi;
        row48[kBlk + i] = a_side ? lo : hi;
        row48[2 * kBlk + i] = a_side ? hi : lo;
    }
}

// ---- BF16 variant (6 partial products) ----------------------------------------------------------------
// A TF32 word carries 11 significant bits, so hi + lo keeps 22 of the 24 bits of an fp32 and the scores sit
// ~4x further from fp64 than the fp32 formula (measured: 1.4e-4 relative at worst over the 1.4e5 models of
// cfg2).  Three BF16 words w0 + w1 + w2 (8 bits each) represent an fp32 EXACTLY, and the six partial products
// of order <= 2 leave 2^-23: fp32-level scores for the same MMA time, because a BF16 MMA step covers K = 16
// in the cycles a TF32 step covers K = 8.  K = 96 = 6 blocks of 16 BF16 = the same 192 bytes per row:
//
//   A = [ w0 | w0 | w0 | w1 | w1 | w2 ],   B = [ w0 | w1 | w2 | w0 | w1 | w0 ]
//
// A 32-bit word of the row holds two consecutive K elements (the lower one in the low half), so word w of a
// row sits at image_index(row, w) in both variants; only the words, the instruction descriptor and the MMA
// kind differ.
constexpr int kK16 = 6 * kBlk;                 // 96

DRB_HD uint16_t bf16_rn(float x) {
#if defined(__CUDA_ARCH__)
    uint16_t h;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)((u >> 16) | ((u & 0xffffu) ? 0x40u : 0u));   // Inf / NaN
    u += 0x7fffu + ((u >> 16) & 1u);           // round to nearest even
    return (uint16_t)(u >> 16);
#endif
}
DRB_HD float bf16_value(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(u);
#else
    memcpy(&f, &u, 4);
#endif
    return f;
}
DRB_HD void bf16_split3(float x, uint16_t* w) {
    w[0] = bf16_rn(x);
    const float r1 = x - bf16_value(w[0]);     // exact
    w[1] = bf16_rn(r1);
    const float r2 = r1 - bf16_value(w[1]);    // exact
    w[2] = bf16_rn(r2);
}
// the 48 32-bit words of one operand row (v = 15 values)
DRB_HD void operand_row_bf16(const float* v, bool a_side, uint32_t* row48w, float v15 = 0.f) {
    uint16_t e[kK16];
    DRB_UNROLL
    for (int i = 0; i < kBlk; ++i) {
        uint16_t w[3] = {0, 0, 0};
        if (i < kFeat) bf16_split3(v[i], w);
        else w[0] = bf16_rn(v15);              // 0, 1, -1 or 1e18: one BF16 word is enough
        // block order of the word index: A (0,0,0,1,1,2), B (0,1,2,0,1,0)
        e[0 * kBlk + i] = w[0];
        e[1 * kBlk + i] = a_side ? w[0] : w[1];
        e[2 * kBlk + i] = a_side ? w[0] : w[2];
        e[3 * kBlk + i] = a_side ? w[1] : w[0];
        e[4 * kBlk + i] = w[1];
        e[5 * kBlk + i] = a_side ? w[2] : w[0];
    }
    DRB_UNROLL
    for (int w = 0; w < kK; ++w) row48w[w] = (uint32_t)e[2 * w] | ((uint32_t)e[2 * w + 1] << 16);
}
// Instruction descriptor of kind::f16 with BF16 operands, fp32 accumulate, K-major A and B
DRB_HD uint32_t instr_desc_bf16() {
    uint32_t d = 0;
    d |= 1u << 4;                              // D format: F32
    d |= 1u << 7;                              // A format: BF16
    d |= 1u << 10;                             // B format: BF16
    d |= (uint32_t)(kTileN >> 3) << 17;
    d |= (uint32_t)(kTileM >> 4) << 24;
    return d;
}
// General form: M x N of the instruction, TF32 or BF16 operands (fp32 accumulate, K-major A and B)
DRB_HD uint32_t instr_desc_mn(int m, int n, bool bf16) {
    const uint32_t fmt = bf16 ? 1u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// TF32 variant as 32-bit words, so both variants share the image writers
DRB_HD void operand_row_words(const float* v, bool a_side, bool bf16, uint32_t* row48w, float v15 = 0.f) {
    if (bf16) {
        operand_row_bf16(v, a_side, row48w, v15);
    } else {
        float row48[kK];
        operand_row(v, a_side, row48, v15);
        DRB_UNROLL
        for (int k = 0; k < kK; ++k) {
#if defined(__CUDA_ARCH__)
            row48w[k] = __float_as_uint(row48[k]);
#else
            memcpy(&row48w[k], &row48[k], 4);
#endif
        }
    }
}

}  // namespace tc
}  // namespace drb
