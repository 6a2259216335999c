// Residual-and-score kernels: one MODEL per thread, the pair's correspondences
// streamed through shared memory by the TMA bulk-copy engine (tile_pipe.cuh).
//
// Replaces scorings/msac_score.py:12-55 (Sampson / soft-MSAC + the arg-max of
// ransac.py:114), model_cl.py:13-26 + loss.py:138-151 (clamped symmetric epipolar
// loss, forward and backward) and rigid_transformation_SVD_based_solver.py:76-89
// (3-D point residual, forward and backward).  The reference materialises
// [M,3,N] / [M,N] temporaries (480 MB at the headline shape); here every thread
// keeps its model in registers, every point is read once from shared memory as a
// warp-wide broadcast, and nothing of size M x N ever exists.
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/drb.h"
#include "drb_common.cuh"
#include "f32x2.cuh"
#include "msac_records.cuh"
#include "rigid_math.cuh"
#include "sampson.cuh"
#include "tile_pipe.cuh"

namespace drb {

constexpr int kScoreThreads = 128;   // episym / rigid kernels: models per CTA
constexpr int kTile = 1024;       // 2-D correspondences (16 B) per stage (episym)
// MSAC: ONE WARP per CTA, 32 models, 6 KB of tiles -> up to 32 CTAs per SM, so the ~4300 live CTAs of
// the headline shape are all resident at once (no second wave, no tail); see profiles/r1_notes.md.
constexpr int kMsacThreads = 32;
constexpr int kMsacTile = 192;
constexpr int kTileRigid = 768;   // 3-D correspondences (24 B) per stage

__global__ void __launch_bounds__(kMsacThreads)
score_msac_kernel(const float* __restrict__ matches, const float* __restrict__ models,
                  const int32_t* __restrict__ count, const int32_t* __restrict__ ids, const float* __restrict__ thr,
                  int M, int N, float* __restrict__ scores, unsigned long long* __restrict__ best_packed) {
    __shared__ __align__(128) float tiles[2 * kMsacTile * 4];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ unsigned long long warp_best[kMsacThreads / 32];
    // blockIdx.x = pair (fastest in dispatch order), blockIdx.y = model block: the live CTAs of every
    // pair are dispatched before the empty tail of the worst-case grid
    const int b = blockIdx.x;
    const int cnt = count ? min(count[b], M) : M;
    const int m0 = blockIdx.y * kMsacThreads;
    if (m0 >= cnt) return;  // whole CTA leaves before any barrier
    const int mi = m0 + threadIdx.x;
    const bool active = mi < cnt;
    float m[9];
    {
        const float* src = models + ((size_t)b * M + (active ? mi : m0)) * 9;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) m[i] = __ldg(src + i);
    }
    const float t = 1.5f * __ldg(thr + b);
    const float inv_thr2 = 1.f / (t * t);

    // model coefficients as packed broadcast operands (ptxas folds the splat into the FFMA2
    // scalar-broadcast operand form, so these cost no extra registers)
    pk2 mp[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) mp[i] = pk2_splat(m[i]);
    pk2 acc = pk2_splat(0.f), accB = pk2_splat(0.f);

    TilePipe<4, kMsacTile> pipe(tiles, bars, matches + (size_t)b * N * 4, N);
    pipe.prologue();
    for (int tI = 0; tI < pipe.n_tiles; ++tI) {
        float4* tile = reinterpret_cast<float4*>(const_cast<float*>(pipe.acquire(tI)));
        const int np = pipe.tile_items(tI);
        // records of two correspondences, padded with NaN to an even count (a full tile is 96 records)
        const int recs_even = msac_interleave(tile, np, threadIdx.x, kMsacThreads);
        __syncthreads();
        if (active) {
            const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(tile);
            for (int i = 0; i < recs_even; i += 2) msac_two_records(t2 + 2 * i, mp, -inv_thr2, acc, accB);
        }
        pipe.release(tI);
    }
    float lo, hi;
    pk2_split(pk2_add(acc, accB), lo, hi);
    const float score = lo + hi;
    if (active && scores) scores[(size_t)b * M + mi] = score;
    unsigned long long key = active ? pack_best(score, ids ? ids[(size_t)b * M + mi] : mi) : 0ull;
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        DRB_UNROLL
        for (int w = 1; w < kMsacThreads / 32; ++w) key = warp_best[w] > key ? warp_best[w] : key;
        if (key) atomicMax(best_packed + b, key);
    }
}

// One CTA per pair: decode the packed arg-max and emit the winner's inlier mask.
__global__ void __launch_bounds__(256)
best_finalize_kernel(const float* __restrict__ matches, const float* __restrict__ models_dense,
                     const unsigned long long* __restrict__ best_packed, const float* __restrict__ thr, int Md, int N,
                     int32_t* __restrict__ best_id, float* __restrict__ best_score, float* __restrict__ best_model,
                     uint8_t* __restrict__ mask, int32_t* __restrict__ ninl) {
    __shared__ int warp_cnt[8];
    const int b = blockIdx.x;
    const unsigned long long key = best_packed[b];
    const int id = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -1;
    float m[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i)
        m[i] = (id >= 0 && id < Md) ? models_dense[((size_t)b * Md + id) * 9 + i] : ((i % 4 == 0) ? 1.f : 0.f);
    if (threadIdx.x == 0) {
        best_id[b] = id;
        best_score[b] = key ? __uint_as_float((unsigned)(key >> 32)) : 0.f;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) best_model[b * 9 + i] = m[i];
    }
    const float t = 1.5f * thr[b];
    const float thr2 = t * t;
    int c = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)b * N + n);
        const Sampson s = sampson(m, p.x, p.y, p.z, p.w);
        const bool in = __fdiv_rn(s.r * s.r, s.j) < thr2;
        if (mask) mask[(size_t)b * N + n] = in ? 1 : 0;
        c += in ? 1 : 0;
    }
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0 && ninl) {
        int tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
        ninl[b] = tot;
    }
}

// ---- symmetric epipolar loss ------------------------------------------------------------------
// ys = r^2 (1/(a + eps) + 1/(c + eps)),  r = x2^T F x1, a = (F x1)_0^2 + (F x1)_1^2,
// c = (F^T x2)_0^2 + (F^T x2)_1^2, eps = 1e-15 (model_cl.py:22-24); loss term min(ys, 1).
template <bool BWD>
__global__ void __launch_bounds__(kScoreThreads)
episym_kernel(const float* __restrict__ pts, const int32_t* __restrict__ npts, const float* __restrict__ models,
              const uint8_t* __restrict__ mvalid, const float* __restrict__ g_row, int K, int P, int n_split, int slice,
              float* __restrict__ out, float* __restrict__ row_out) {
    __shared__ __align__(128) float tiles[2 * kTile * 4];
    __shared__ __align__(8) uint64_t bars[2];
    const int b = blockIdx.y;
    const int k = blockIdx.x * kScoreThreads + threadIdx.x;
    const bool active = k < K && (!mvalid || mvalid[(size_t)b * K + k]);
    const int np_total = npts ? min(npts[b], P) : P;
    // This CTA's slice of the points.  The pairs of a batch bring different numbers of points (the GT inliers: 400 to
    // 1200 at cfg3 / cfg5), so a pair is cut into as many EQUAL slices as it needs to stay under `slice` points, not
    // into a fixed number: every CTA of the grid then has about the same work and the tail of the launch is not the
    // longest pairs alone (loss forward + backward at cfg5: 0.118 -> 0.105 ms).  CTAs past a pair's count leave.
    const int n_mine = n_split == 1 ? 1 : min(n_split, max(1, (np_total + slice - 1) / slice));
    if ((int)blockIdx.z >= n_mine) return;
    const int per = ((np_total + n_mine - 1) / n_mine + 3) & ~3;
    const int p_begin = min(np_total, (int)blockIdx.z * per);
    const int p_end = min(np_total, p_begin + per);
    float m[9];
    {
        const float* src = models + ((size_t)b * K + (k < K ? k : 0)) * 9;
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) m[i] = __ldg(src + i);
    }
    float acc = 0.f;
    float g[9];
    DRB_UNROLL
    for (int i = 0; i < 9; ++i) g[i] = 0.f;
    const float eps = 1e-15f;

    TilePipe<4, kTile> pipe(tiles, bars, pts + ((size_t)b * P + p_begin) * 4, p_end - p_begin);
    pipe.prologue();
    for (int tI = 0; tI < pipe.n_tiles; ++tI) {
        const float4* tile = reinterpret_cast<const float4*>(pipe.acquire(tI));
        const int np = pipe.tile_items(tI);
        if (active) {
#pragma unroll 2
            for (int i = 0; i < np; ++i) {
                const float4 p = tile[i];
                const float x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
                const float e0 = fmaf(m[0], x1, fmaf(m[1], y1, m[2]));
                const float e1 = fmaf(m[3], x1, fmaf(m[4], y1, m[5]));
                const float e2 = fmaf(m[6], x1, fmaf(m[7], y1, m[8]));
                const float f0 = fmaf(m[0], x2, fmaf(m[3], y2, m[6]));
                const float f1 = fmaf(m[1], x2, fmaf(m[4], y2, m[7]));
                const float r = fmaf(x2, e0, fmaf(y2, e1, e2));
                const float a = fmaf(e0, e0, e1 * e1) + eps;
                const float c = fmaf(f0, f0, f1 * f1) + eps;
                const float ia = rcp_approx(a), ic = rcp_approx(c);
                const float w = ia + ic;
                const float ys = r * r * w;
                if (!BWD) {
                    acc += fminf(ys, 1.f);
                } else if (!(ys < 1.f)) {
                    acc += 1.f;                      // clamped term: value 1, no gradient (NaN counts as clamped)
                } else {
                    acc += ys;
                    // d ys = 2 r w dr - r^2 (ia^2 da + ic^2 dc)
                    const float cr = 2.f * r * w;
                    const float r2 = r * r;
                    const float ca = -r2 * ia * ia, cc = -r2 * ic * ic;
                    // dr/dF_ij = x2_i x1_j ; da/dF = 2 e0 d e0 + 2 e1 d e1 ; dc/dF = 2 f0 d f0 + 2 f1 d f1
                    const float a0 = 2.f * ca * e0, a1 = 2.f * ca * e1;
                    const float c0 = 2.f * cc * f0, c1 = 2.f * cc * f1;
                    g[0] += cr * x2 * x1 + a0 * x1 + c0 * x2;
                    g[1] += cr * x2 * y1 + a0 * y1 + c1 * x2;
                    g[2] += cr * x2 + a0;
                    g[3] += cr * y2 * x1 + a1 * x1 + c0 * y2;
                    g[4] += cr * y2 * y1 + a1 * y1 + c1 * y2;
                    g[5] += cr * y2 + a1;
                    g[6] += cr * x1 + c0;
                    g[7] += cr * y1 + c1;
                    g[8] += cr;
                }
            }
        }
        pipe.release(tI);
    }
    if (!active) return;
    if (!BWD) {
        if (n_split == 1) out[(size_t)b * K + k] = acc;
        else atomicAdd(out + (size_t)b * K + k, acc);
    } else {
        const float gr = g_row[(size_t)b * K + k];
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) {
            if (n_split == 1) out[((size_t)b * K + k) * 9 + i] = g[i] * gr;
            else atomicAdd(out + ((size_t)b * K + k) * 9 + i, g[i] * gr);
        }
        if (row_out) {                               // fused forward + backward: the row sums come out of the same pass
            if (n_split == 1) row_out[(size_t)b * K + k] = acc;
            else atomicAdd(row_out + (size_t)b * K + k, acc);
        }
    }
}

// ---- rigid residual ---------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kScoreThreads)
rigid_residual_kernel(const float* __restrict__ points, const float* __restrict__ models,
                      const float* __restrict__ g_res, int K, int N, int n_split, float threshold,
                      float* __restrict__ out, int32_t* __restrict__ ninl, float* __restrict__ res_out) {
    __shared__ __align__(128) float tiles[2 * kTileRigid * 6];
    __shared__ __align__(8) uint64_t bars[2];
    const int b = blockIdx.y;
    const int k = blockIdx.x * kScoreThreads + threadIdx.x;
    const bool active = k < K;
    const int per = ((N + n_split - 1) / n_split + 3) & ~3;
    const int p_begin = min(N, (int)blockIdx.z * per);
    const int p_end = min(N, p_begin + per);
    float m[12];
    {
        const float* src = models + ((size_t)b * K + (active ? k : 0)) * 16;
        DRB_UNROLL
        for (int i = 0; i < 12; ++i) m[i] = __ldg(src + i);
    }
    float acc = 0.f;
    int cnt = 0;
    float g[12];
    DRB_UNROLL
    for (int i = 0; i < 12; ++i) g[i] = 0.f;

    TilePipe<6, kTileRigid> pipe(tiles, bars, points + ((size_t)b * N + p_begin) * 6, p_end - p_begin);
    pipe.prologue();
    for (int tI = 0; tI < pipe.n_tiles; ++tI) {
        const float2* tile = reinterpret_cast<const float2*>(pipe.acquire(tI));
        const int np = pipe.tile_items(tI);
        if (active) {
#pragma unroll 2
            for (int i = 0; i < np; ++i) {
                const float2 a = tile[3 * i], c = tile[3 * i + 1], e = tile[3 * i + 2];
                const float px = a.x, py = a.y, pz = c.x, qx = c.y, qy = e.x, qz = e.y;
                const float dx = qx - fmaf(m[0], px, fmaf(m[1], py, fmaf(m[2], pz, m[3])));
                const float dy = qy - fmaf(m[4], px, fmaf(m[5], py, fmaf(m[6], pz, m[7])));
                const float dz = qz - fmaf(m[8], px, fmaf(m[9], py, fmaf(m[10], pz, m[11])));
                if (!BWD) {
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    acc += d2;
                    cnt += d2 < threshold ? 1 : 0;
                } else {
                    acc = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, acc)));
                    g[0] -= dx * px; g[1] -= dx * py; g[2] -= dx * pz; g[3] -= dx;
                    g[4] -= dy * px; g[5] -= dy * py; g[6] -= dy * pz; g[7] -= dy;
                    g[8] -= dz * px; g[9] -= dz * py; g[10] -= dz * pz; g[11] -= dz;
                }
            }
        }
        pipe.release(tI);
    }
    if (!active) return;
    if (!BWD) {
        if (n_split == 1) {
            out[(size_t)b * K + k] = acc;
            if (ninl) ninl[(size_t)b * K + k] = cnt;
        } else {
            atomicAdd(out + (size_t)b * K + k, acc);
            if (ninl) atomicAdd(ninl + (size_t)b * K + k, cnt);
        }
    } else {
        const float gr = 2.f * g_res[(size_t)b * K + k];
        DRB_UNROLL
        for (int i = 0; i < 12; ++i) {
            if (n_split == 1) out[((size_t)b * K + k) * 16 + i] = g[i] * gr;
            else atomicAdd(out + ((size_t)b * K + k) * 16 + i, g[i] * gr);
        }
        if (res_out) {
            if (n_split == 1) res_out[(size_t)b * K + k] = acc;
            else atomicAdd(res_out + (size_t)b * K + k, acc);
        }
    }
}


// ---- rigid residual from the points' second moments -------------------------------------------------------------
// sum_n ||q_n - (R p_n + t)||^2 and its gradient in (R, t) are quadratic forms in the model whose coefficients are
// sums over the points ALONE:
//     S_pp = sum p p',  S_qp = sum q p',  s_p = sum p,  s_q = sum q,  s_qq = sum |q|^2          (22 numbers)
//     sum d_i p_j = S_qp[i][j] - R_i . S_pp[:, j] - t_i s_p[j],      sum d_i = s_q[i] - R_i . s_p - N t_i
//     sum |d|^2   = s_qq - 2 sum_i (R_i . S_qp[i] + t_i s_q[i]) + sum_i (R_i S_pp R_i' + 2 t_i R_i . s_p + N t_i^2)
// so a pair costs O(N + K) instead of the O(N K) of rigid_residual_kernel (cfg4, N = 50 000, K = 1 000: 1.05 ms ->
// tens of microseconds).  The moments are accumulated in fp64 from exact products of the fp32 coordinates (the sums
// cancel for a model that fits: 5e4 points of |q|^2 ~ 3 against a residual sum of ~15) and every CTA recomputes its
// pair's 22 moments -- 1.2 MB from L2 at cfg4 -- rather than taking a workspace through the ABI.  Same results as the
// per-point kernel to fp32 rounding (closer to the reference run in fp64); the per-point kernel stays for callers
// that want the inlier counts, which are not a function of the moments.
constexpr int kMomThreads = 256;
constexpr int kMoments = kRigidMoments;

template <bool BWD>
__global__ void __launch_bounds__(kMomThreads)
rigid_residual_moments_kernel(const float* __restrict__ points, const float* __restrict__ models,
                              const float* __restrict__ g_res, int K, int N, float* __restrict__ g_models,
                              float* __restrict__ res_out) {
    __shared__ double red[kMomThreads / 32][kMoments];
    __shared__ double mom[kMoments];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s[kMoments];
    DRB_UNROLL
    for (int i = 0; i < kMoments; ++i) s[i] = 0.0;
    const float2* src = reinterpret_cast<const float2*>(points + (size_t)b * N * 6);
    for (int n = threadIdx.x; n < N; n += kMomThreads) {
        const float2 a = __ldg(src + 3 * n), c = __ldg(src + 3 * n + 1), e = __ldg(src + 3 * n + 2);
        const double p[3] = {(double)a.x, (double)a.y, (double)c.x}, q[3] = {(double)c.y, (double)e.x, (double)e.y};
        rigid_moments_add<double>(p, q, s);
    }
    DRB_UNROLL
    for (int i = 0; i < kMoments; ++i) {
        DRB_UNROLL
        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    if (lane == 0) {
        DRB_UNROLL
        for (int i = 0; i < kMoments; ++i) red[warp][i] = s[i];
    }
    __syncthreads();
    if (threadIdx.x < kMoments) {
        double t = 0.0;
        DRB_UNROLL
        for (int w = 0; w < kMomThreads / 32; ++w) t += red[w][threadIdx.x];
        mom[threadIdx.x] = t;
    }
    __syncthreads();
    const int k = blockIdx.x * kMomThreads + threadIdx.x;
    if (k >= K) return;
    float m12[12];
    DRB_UNROLL
    for (int i = 0; i < 12; ++i) m12[i] = __ldg(models + ((size_t)b * K + k) * 16 + i);
    double res, g[12];
    rigid_residual_from_moments<float, double>(mom, (double)N, m12, res, g);
    if (res_out) res_out[(size_t)b * K + k] = (float)res;
    if (BWD) {
        const double gr = 2.0 * (double)__ldg(g_res + (size_t)b * K + k);
        DRB_UNROLL
        for (int i = 0; i < 12; ++i) g_models[((size_t)b * K + k) * 16 + i] = (float)(g[i] * gr);
    }
}

// measurement switch: DRB_RIGID_RESIDUAL=points keeps the O(N K) kernel for the calls that do not need it
static bool rigid_by_moments() {
    static const bool on = []() {
        const char* e = getenv("DRB_RIGID_RESIDUAL");
        return !(e != nullptr && e[0] == 'p');
    }();
    return on;
}

// episym launches: slices of about `slice` points (>= 256: every slice costs ten atomics per model), as many as it
// takes to put ~16 CTAs on every SM; returns the grid's z extent (the longest pair's slice count)
static int pick_slices(int ctas, int n_items, int& slice) {
    const long long want = 148LL * 16;
    slice = 256;
    const long long at256 = (long long)ctas * ((n_items + 255) / 256);
    if (at256 > want) slice = (int)(((long long)n_items * ctas + want - 1) / want);
    slice = (slice + 3) & ~3;
    if (slice < 256) slice = 256;
    int n = (n_items + slice - 1) / slice;
    if (n < 1) n = 1;
    if (n > 65535) n = 65535;
    return n;
}

static int pick_split(int ctas, int n_items) {
    // enough CTAs to cover 148 SMs a few times over, but never slices thinner than a tile
    int split = (148 * 4 + ctas - 1) / ctas;
    const int max_split = (n_items + kTile - 1) / kTile;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    return split;
}

}  // namespace drb

using namespace drb;

#define DRB_CHECK_LAUNCH() return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA

extern "C" int drb_score_msac(const float* matches, const float* models, const int32_t* count, const int32_t* ids,
                              const float* thr, int B, int M, int N, float* scores,
                              unsigned long long* best_packed, void* stream) {
    if (!matches || !models || !thr || !best_packed) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || M <= 0 || N <= 0 || (M + kMsacThreads - 1) / kMsacThreads > 65535) return DRB_ERR_BAD_SHAPE;
    dim3 grid(B, (M + kMsacThreads - 1) / kMsacThreads);
    score_msac_kernel<<<grid, kMsacThreads, 0, (cudaStream_t)stream>>>(matches, models, count, ids, thr, M, N, scores,
                                                                       best_packed);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_best_finalize(const float* matches, const float* models_dense,
                                 const unsigned long long* best_packed, const float* thr, int B, int Md, int N,
                                 int32_t* best_id, float* best_score, float* best_model, uint8_t* mask, int32_t* ninl,
                                 void* stream) {
    if (!matches || !models_dense || !best_packed || !thr || !best_id || !best_score || !best_model)
        return DRB_ERR_NULL_POINTER;
    if (B <= 0 || Md <= 0 || N <= 0) return DRB_ERR_BAD_SHAPE;
    best_finalize_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(matches, models_dense, best_packed, thr, Md, N, best_id,
                                                              best_score, best_model, mask, ninl);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_episym_forward(const float* pts, const int32_t* npts, const float* models, const uint8_t* mvalid,
                                  int B, int K, int P, float* row_sum, void* stream) {
    if (!pts || !models || !row_sum) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || P <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    int slice;
    const int split = pick_slices(gx * B, P, slice);
    // inactive (invalid) models keep a zero row sum: callers multiply by the validity flag, and 0 * garbage may be NaN
    if (split > 1 || mvalid) cudaMemsetAsync(row_sum, 0, sizeof(float) * (size_t)B * K, (cudaStream_t)stream);
    episym_kernel<false><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(pts, npts, models, mvalid,
                                                                                       nullptr, K, P, split, slice,
                                                                                       row_sum, nullptr);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_episym_backward(const float* pts, const int32_t* npts, const float* models, const uint8_t* mvalid,
                                   const float* g_row, int B, int K, int P, float* g_models, void* stream) {
    if (!pts || !models || !g_row || !g_models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || P <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    int slice;
    const int split = pick_slices(gx * B, P, slice);
    // inactive (invalid) models keep a zero gradient
    cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 9, (cudaStream_t)stream);
    episym_kernel<true><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(pts, npts, models, mvalid, g_row,
                                                                                      K, P, split, slice, g_models,
                                                                                      nullptr);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_episym_forward_backward(const float* pts, const int32_t* npts, const float* models,
                                           const uint8_t* mvalid, const float* g_row, int B, int K, int P,
                                           float* row_sum, float* g_models, void* stream) {
    if (!pts || !models || !g_row || !row_sum || !g_models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || P <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    int slice;
    const int split = pick_slices(gx * B, P, slice);
    cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 9, (cudaStream_t)stream);
    cudaMemsetAsync(row_sum, 0, sizeof(float) * (size_t)B * K, (cudaStream_t)stream);   // inactive models: 0
    episym_kernel<true><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(pts, npts, models, mvalid, g_row,
                                                                                      K, P, split, slice, g_models,
                                                                                      row_sum);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_rigid_residual_forward(const float* points, const float* models, int B, int K, int N,
                                          float threshold, float* res_sum, int32_t* ninl, void* stream) {
    if (!points || !models || !res_sum) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    if (!ninl && rigid_by_moments()) {
        rigid_residual_moments_kernel<false><<<dim3((K + kMomThreads - 1) / kMomThreads, B), kMomThreads, 0,
                                               (cudaStream_t)stream>>>(points, models, nullptr, K, N, nullptr, res_sum);
        DRB_CHECK_LAUNCH();
    }
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    const int split = pick_split(gx * B, N);
    if (split > 1) {
        cudaMemsetAsync(res_sum, 0, sizeof(float) * (size_t)B * K, (cudaStream_t)stream);
        if (ninl) cudaMemsetAsync(ninl, 0, sizeof(int32_t) * (size_t)B * K, (cudaStream_t)stream);
    }
    rigid_residual_kernel<false><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(
        points, models, nullptr, K, N, split, threshold, res_sum, ninl, nullptr);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_rigid_residual_backward(const float* points, const float* models, const float* g_res, int B, int K,
                                           int N, float* g_models, void* stream) {
    if (!points || !models || !g_res || !g_models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    if (rigid_by_moments()) {
        cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 16, (cudaStream_t)stream);   // the fourth rows
        rigid_residual_moments_kernel<true><<<dim3((K + kMomThreads - 1) / kMomThreads, B), kMomThreads, 0,
                                              (cudaStream_t)stream>>>(points, models, g_res, K, N, g_models, nullptr);
        DRB_CHECK_LAUNCH();
    }
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    const int split = pick_split(gx * B, N);
    cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 16, (cudaStream_t)stream);
    rigid_residual_kernel<true><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(
        points, models, g_res, K, N, split, 0.f, g_models, nullptr, nullptr);
    DRB_CHECK_LAUNCH();
}

extern "C" int drb_rigid_residual_forward_backward(const float* points, const float* models, const float* g_res, int B,
                                                   int K, int N, float* res_sum, float* g_models, void* stream) {
    if (!points || !models || !g_res || !res_sum || !g_models) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || B > 65535) return DRB_ERR_BAD_SHAPE;
    if (rigid_by_moments()) {
        cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 16, (cudaStream_t)stream);   // the fourth rows
        rigid_residual_moments_kernel<true><<<dim3((K + kMomThreads - 1) / kMomThreads, B), kMomThreads, 0,
                                              (cudaStream_t)stream>>>(points, models, g_res, K, N, g_models, res_sum);
        DRB_CHECK_LAUNCH();
    }
    const int gx = (K + kScoreThreads - 1) / kScoreThreads;
    const int split = pick_split(gx * B, N);
    cudaMemsetAsync(g_models, 0, sizeof(float) * (size_t)B * K * 16, (cudaStream_t)stream);
    if (split > 1) cudaMemsetAsync(res_sum, 0, sizeof(float) * (size_t)B * K, (cudaStream_t)stream);
    rigid_residual_kernel<true><<<dim3(gx, B, split), kScoreThreads, 0, (cudaStream_t)stream>>>(
        points, models, g_res, K, N, split, 0.f, g_models, nullptr, res_sum);
    DRB_CHECK_LAUNCH();
}
