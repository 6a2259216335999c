// Gumbel top-s sampler: one warp per (pair, hypothesis).
//
// Replaces GumbelSoftmaxSampler.sample (samplers/gumbel_sampler.py:25-42) and the
// boolean-mask gather of ransac.py:64-65.  The reference materialises five K x N
// temporaries (repeat, gumbels, softmax, one-hot, straight-through) and a K x N x D
// product; here a warp streams the N keys once, keeps the running top-s in the
// registers of lanes 0..s-1 and an online log-sum-exp, and writes s indices.
//
// Noise: injected ([B,K,N], parity mode -- keys are formed with IEEE add/div so the
// top-s matches torch.topk bit for bit) or Philox4x32-10 generated in-kernel.
#include <cuda_runtime.h>

#include "../../include/drb.h"
#include "device_cfg.cuh"
#include "drb_common.cuh"
#include "philox.cuh"

namespace drb {

constexpr int kSamplerWarps = 8;

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct KeyQuad {
    float k[4];
};

// Four consecutive keys n0..n0+3 of hypothesis row (b, k); entries past N are -inf.
__device__ __forceinline__ KeyQuad load_keys(const float* __restrict__ logits_b, const float* __restrict__ noise_row,
                                             float* __restrict__ noise_out_row, bool vec, int n0, int N, float tau,
                                             uint32_t k, uint32_t b, uint64_t seed, uint64_t offset) {
    float l[4], g[4];
    if (vec) {
        const float4 lv = __ldg(reinterpret_cast<const float4*>(logits_b + n0));
        l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
    } else {
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) l[i] = (n0 + i < N) ? __ldg(logits_b + n0 + i) : 0.f;
    }
    if (noise_row != nullptr) {
        if (vec) {
            const float4 gv = __ldcs(reinterpret_cast<const float4*>(noise_row + n0));
            g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
        } else {
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) g[i] = (n0 + i < N) ? __ldcs(noise_row + n0 + i) : 0.f;
        }
    } else {
        const Philox4 r = philox4x32_10((uint32_t)(n0 >> 2), k, b, (uint32_t)offset, (uint32_t)seed,
                                        (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32));
        g[0] = gumbel_from_bits(r.x);
        g[1] = gumbel_from_bits(r.y);
        g[2] = gumbel_from_bits(r.z);
        g[3] = gumbel_from_bits(r.w);
    }
    if (noise_out_row != nullptr) {
        DRB_UNROLL
        for (int i = 0; i < 4; ++i)
            if (n0 + i < N) noise_out_row[n0 + i] = g[i];
    }
    KeyQuad q;
    if (tau == 1.0f) {  // x / 1 == x exactly: skip the IEEE division (warp-uniform branch)
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) q.k[i] = (n0 + i < N) ? __fadd_rn(l[i], g[i]) : -INFINITY;
    } else {
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) q.k[i] = (n0 + i < N) ? __fdiv_rn(__fadd_rn(l[i], g[i]), tau) : -INFINITY;
    }
    return q;
}

// Warp-shared running top-S (element j in lane j, descending).  `cand_v`/`cand_i` are this
// lane's candidate; every lane whose candidate beats the S-th best is merged in.
template <int S>
__device__ __forceinline__ void merge_candidates(float cand_v, int cand_i, float& top_v, int& top_i, float& thr,
                                                 int lane) {
    const unsigned FULL = 0xffffffffu;
    unsigned cand = __ballot_sync(FULL, cand_v > thr);
    while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        const float cv = __shfl_sync(FULL, cand_v, src);
        if (!(cv > thr)) continue;  // threshold moved since the ballot (warp-uniform)
        const int ci = __shfl_sync(FULL, cand_i, src);
        const float up_v = __shfl_up_sync(FULL, top_v, 1);
        const int up_i = __shfl_up_sync(FULL, top_i, 1);
        if (cv > top_v) {
            if (lane > 0 && cv > up_v) {
                top_v = up_v;
                top_i = up_i;
            } else {
                top_v = cv;
                top_i = ci;
            }
        }
        thr = __shfl_sync(FULL, top_v, S - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path (no injected noise, no log-sum-exp wanted = test mode): the Gumbel-max trick as an
// exponential race.  key = logit + G, G = -ln(-ln v)  <=>  rank by  log2(v) * exp(-logit)
// (largest wins), which needs ONE SFU op per element instead of two logarithms and a division;
// exp(-logit[n]) is tabulated once per CTA in shared memory.  tau > 0 does not change the
// ranking.  Same Philox stream and the same v as the exact kernel, so both select the same
// points (up to ties at the last ulp; tests/test_gpu_parity.py::test_sampler_race_equals_exact).
constexpr int kRaceMaxN = 8192;

template <int S>
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_race_kernel(const float* __restrict__ logits, uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ offset_dev, int K,
                   int N, int32_t* __restrict__ idx_out) {
    if (offset_dev) offset += *offset_dev;  // stream position kept on the device (CUDA-graph replay draws fresh noise)
    __shared__ __align__(16) float winv[kRaceMaxN];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int k = blockIdx.x * kSamplerWarps + warp;
    const float* logits_b = logits + (size_t)b * N;
    const int n_pad = ((N + 127) / 128) * 128;
    for (int n = threadIdx.x; n < n_pad; n += blockDim.x)
        winv[n] = n < N ? __expf(-__ldg(logits_b + n)) : INFINITY;  // lg2(v) < 0: padded entries rank -inf
    __syncthreads();
    if (k >= K) return;
    const long long row = (long long)b * K + k;
    float top_v = -INFINITY, thr = -INFINITY;
    int top_i = -1;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
    for (int n0 = lane * 4; n0 < n_pad; n0 += 128) {
        const Philox4 r = philox4x32_10((uint32_t)(n0 >> 2), (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0, k1);
        const float4 w = *reinterpret_cast<const float4*>(winv + n0);
        float t[4];
        t[0] = lg2_approx(uniform_from_bits(r.x)) * w.x;
        t[1] = lg2_approx(uniform_from_bits(r.y)) * w.y;
        t[2] = lg2_approx(uniform_from_bits(r.z)) * w.z;
        t[3] = lg2_approx(uniform_from_bits(r.w)) * w.w;
        // lane-local best of the four first: one ballot per iteration in the common case
        float bv = t[0];
        int bi = 0;
        DRB_UNROLL
        for (int i = 1; i < 4; ++i)
            if (t[i] > bv) { bv = t[i]; bi = i; }
        if (n0 < 128) {
            // first sweep: the list is empty, so every lane would be a candidate.  Take the S best
            // of the 32 lane-maxima by S rounds of warp arg-max instead (straight-line code).
            float cur = bv;
            DRB_UNROLL
            for (int j = 0; j < S; ++j) {
                float m = cur;
                DRB_UNROLL
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const int owner = __ffs(__ballot_sync(0xffffffffu, cur == m)) - 1;
                const int mi = __shfl_sync(0xffffffffu, n0 + bi, owner);
                if (lane == j) { top_v = m; top_i = mi; }
                if (lane == owner) cur = -INFINITY;
            }
            thr = __shfl_sync(0xffffffffu, top_v, S - 1);
        } else {
            merge_candidates<S>(bv, n0 + bi, top_v, top_i, thr, lane);
        }
        // rare: another key of the same lane also beats the (updated) threshold
        float second = -INFINITY;
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) second = (i == bi) ? second : fmaxf(second, t[i]);
        if (__any_sync(0xffffffffu, second > thr)) {
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) merge_candidates<S>((i == bi) ? -INFINITY : t[i], n0 + i, top_v, top_i, thr, lane);
        }
    }
    int rank = 0;
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        const int oj = __shfl_sync(0xffffffffu, top_i, j);
        rank += (oj < top_i) ? 1 : 0;
    }
    if (lane < S) idx_out[(size_t)row * S + rank] = top_i;
}

// ---------------------------------------------------------------------------------------------
// Set sampler (test mode, nothing but the index sets is wanted).  Taking the s largest of
// logit + Gumbel noise IS sampling s items without replacement from softmax(logits) (the
// Plackett-Luce law behind the Gumbel-top-k trick), so the sets can be drawn directly: one THREAD
// per hypothesis makes s inverse-CDF draws on a per-pair prefix sum held in shared memory and
// rejects repeats.  O(s log N) per hypothesis instead of O(N): the K x N noise the reference
// (and the kernels above) generate only to throw away is never produced.  The distribution is
// identical; the realisation differs, which is immaterial without an injected-noise reference.
constexpr int kSetThreads = 256;

template <int S>
__global__ void __launch_bounds__(kSetThreads)
sample_sets_kernel(const float* __restrict__ logits, uint64_t seed, uint64_t offset,
                   const unsigned long long* __restrict__ offset_dev, int K, int N, int32_t* __restrict__ idx_out) {
    extern __shared__ float cdf[];  // N inclusive prefix sums of exp(logit - max)
    if (offset_dev) offset += *offset_dev;  // stream position kept on the device: a captured graph replays with fresh draws
    __shared__ float red[kSetThreads / 32];
    __shared__ float warp_tot[kSetThreads / 32];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* lg = logits + (size_t)b * N;
    // 1. max logit (so that exp cannot overflow)
    float mx = -INFINITY;
    for (int n = tid; n < N; n += kSetThreads) mx = fmaxf(mx, __ldg(lg + n));
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    DRB_UNROLL
    for (int w = 1; w < kSetThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    // 2. block-wide inclusive scan of the weights, each thread owning a contiguous chunk
    const int chunk = (N + kSetThreads - 1) / kSetThreads;
    const int n_lo = min(N, tid * chunk), n_hi = min(N, n_lo + chunk);
    float local = 0.f;
    for (int n = n_lo; n < n_hi; ++n) {
        local += __expf(__ldg(lg + n) - mx);
        cdf[n] = local;
    }
    float incl = local;
    DRB_UNROLL
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    float base = incl - local;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    for (int n = n_lo; n < n_hi; ++n) cdf[n] += base;
    __syncthreads();
    const int k = blockIdx.x * kSetThreads + tid;
    if (k >= K) return;
    const float total = cdf[N - 1];
    int chosen[S];
    int count = 0;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
    for (int call = 0; call < 16 && count < S; ++call) {
        const Philox4 r = philox4x32_10((uint32_t)call, (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0, k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        DRB_UNROLL
        for (int t = 0; t < 4; ++t) {
            if (count < S) {
                // 24 random bits -> u in [0, total); first n with cdf[n] > u
                const float u = (float)(rr[t] >> 8) * 5.9604644775390625e-08f * total;
                int lo = 0, hi = N - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
                }
                bool dup = false;
                DRB_UNROLL
                for (int j = 0; j < S; ++j) dup = dup || (j < count && chosen[j] == lo);
                if (!dup) {
                    DRB_UNROLL
                    for (int j = 0; j < S; ++j)
                        if (j == count) chosen[j] = lo;
                    ++count;
                }
            }
        }
    }
    // Peaked weights (almost all the mass on fewer than S items): 64 rejected draws.  The remaining members are
    // then drawn EXACTLY -- Gumbel-max over the items not chosen yet, keys logit + G on fresh Philox words, the
    // `S - count` largest win -- which is the same Plackett-Luce law continued, and resolves weights that the fp32
    // prefix sums cannot (a weight of 1e-10 beside a total of 1).  O(N) for this thread; rare by construction.
    if (count < S) {
        const int need = S - count;
        float bk[S];
        int bi[S];
        DRB_UNROLL
        for (int j = 0; j < S; ++j) { bk[j] = -INFINITY; bi[j] = -1; }
        for (int n4 = 0; n4 < N; n4 += 4) {
            const Philox4 r = philox4x32_10((uint32_t)(16 + (n4 >> 2)), (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0, k1);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            DRB_UNROLL
            for (int t = 0; t < 4; ++t) {
                const int n = n4 + t;
                if (n >= N) break;
                bool dup = false;
                DRB_UNROLL
                for (int j = 0; j < S; ++j) dup = dup || (j < count && chosen[j] == n);
                if (dup) continue;
                const float u = ((float)(rr[t] >> 8) + 0.5f) * 5.9604644775390625e-08f;
                float key = __ldg(lg + n) - __logf(-__logf(u));
                int id = n;
                // insertion into the descending list of the `need` best (ties: lower index first, like top-k)
                DRB_UNROLL
                for (int j = 0; j < S; ++j) {
                    if (j < need && (key > bk[j] || (key == bk[j] && bi[j] < 0))) {
                        const float tk = bk[j]; const int ti = bi[j];
                        bk[j] = key; bi[j] = id;
                        key = tk; id = ti;
                    }
                }
            }
        }
        DRB_UNROLL
        for (int q = 0; q < S; ++q) {
            if (q < need && bi[q] >= 0) {
                DRB_UNROLL
                for (int j = 0; j < S; ++j)
                    if (j == count) chosen[j] = bi[q];
                ++count;
            }
        }
    }
    // ascending order, like the reference's boolean-mask gather (ransac.py:65)
    DRB_UNROLL
    for (int i = 1; i < S; ++i) {
        DRB_UNROLL
        for (int j = i; j > 0; --j) {
            if (chosen[j - 1] > chosen[j]) { const int t = chosen[j]; chosen[j] = chosen[j - 1]; chosen[j - 1] = t; }
        }
    }
    const long long row = (long long)b * K + k;
    DRB_UNROLL
    for (int j = 0; j < S; ++j) idx_out[row * S + j] = chosen[j];
}

// ---------------------------------------------------------------------------------------------
// Training forward, in-kernel noise, tau = 1: indices + log-sum-exp + selected keys, with the race
// formulation for the ranking (rank by t = log2(u) * exp(-(l - lmax)), largest wins) and
//   sum_n exp(l_n + g_n) = exp(lmax) * sum_n w_n / e_n,   w = exp(l - lmax),  e = -ln u,
// for the normaliser: one log2 and one reciprocal per element instead of two logs, an IEEE add and
// an online softmax.  Draws the SAME noise as sample_kernel (same Philox counters, same clamp), so
// the backward (which regenerates it) and the exact kernel agree with it to rounding.
//
// Top-s without a running top-s: P(t_n > x) = 1 - 2^(x w_n) ~ -x w_n ln 2, so the number of elements above
// x0 = -3s / (W ln 2), W = sum_n w_n, is Poisson with mean ~3s (fewer than s -- a second sweep -- with probability
// 0.6 % at s = 3, 0.09 % at s = 5; the mean was 2s until round 2: 6.2 % / 2.9 %).  The sweep only APPENDS those few to a
// per-hypothesis list in shared memory; the s largest of the list are ranked after the sweep.  Fewer than s
// (or more than the list holds) sends the hypothesis through a second, unfiltered sweep with the warp-shared
// running top-s (a few percent of the hypotheses for reference-like weights; every one in the worst case).
// N is swept in chunks of kTrainChunk correspondences whose tables live in shared memory, so any N runs here.
// Hypotheses (warps) per CTA share one pair of tables.  8 when N fits one chunk (more, smaller CTAs; measured
// faster at N = 2000), 16 when the tables are rebuilt chunk after chunk (halves the rebuilding per hypothesis).
// Chunk = what keeps two float tables + the candidate lists under 48 KB of static shared memory.
constexpr int kTrainCand = 64;
constexpr int train_chunk(int warps) { return warps == 8 ? 5376 : 4992; }

template <int S, bool FILTER>
__device__ __forceinline__ void train_sweep(const float* __restrict__ wtab, const float* __restrict__ winv, int base,
                                            int len_pad, uint32_t k, uint32_t b, uint32_t c3, uint32_t k0, uint32_t k1,
                                            float x0, float* cand_v, int* cand_i, int* cand_n, float& zacc, float& top_v,
                                            int& top_i, float& thr, int lane) {
    const unsigned FULL = 0xffffffffu;
    for (int n0 = lane * 4; n0 < len_pad; n0 += 128) {
        const Philox4 r = philox4x32_10((uint32_t)((base + n0) >> 2), k, b, c3, k0, k1);
        const float4 wi = *reinterpret_cast<const float4*>(winv + n0);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        const float wiv[4] = {wi.x, wi.y, wi.z, wi.w};
        float t[4];
        if (FILTER) {
            const float4 ww = *reinterpret_cast<const float4*>(wtab + n0);
            const float wwv[4] = {ww.x, ww.y, ww.z, ww.w};
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) {
                const float lg = fminf(lg2_approx(uniform_from_bits(rr[i])), -1.4426950408889634e-10f);
                t[i] = lg * wiv[i];
                zacc = fmaf(wwv[i], rcp_fast(lg), zacc);  // sum of w / log2(u)  (negative)
            }
            const float m4 = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
            if (m4 > x0) {
                DRB_UNROLL
                for (int i = 0; i < 4; ++i)
                    if (t[i] > x0) {
                        const int pos = atomicAdd(cand_n, 1);
                        if (pos < kTrainCand) {
                            cand_v[pos] = t[i];
                            cand_i[pos] = base + n0 + i;
                        }
                    }
            }
        } else {
            DRB_UNROLL
            for (int i = 0; i < 4; ++i)
                t[i] = fminf(lg2_approx(uniform_from_bits(rr[i])), -1.4426950408889634e-10f) * wiv[i];
            float bv = t[0];
            int bi = 0;
            DRB_UNROLL
            for (int i = 1; i < 4; ++i)
                if (t[i] > bv) { bv = t[i]; bi = i; }
            merge_candidates<S>(bv, base + n0 + bi, top_v, top_i, thr, lane);
            float second = -INFINITY;
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) second = (i == bi) ? second : fmaxf(second, t[i]);
            if (__any_sync(FULL, second > thr)) {
                DRB_UNROLL
                for (int i = 0; i < 4; ++i)
                    merge_candidates<S>((i == bi) ? -INFINITY : t[i], base + n0 + i, top_v, top_i, thr, lane);
            }
        }
    }
}

template <int S, int kTrainWarps>
__global__ void __launch_bounds__(kTrainWarps * 32)
sample_train_kernel(const float* __restrict__ logits, uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ offset_dev, int K,
                    int N, int32_t* __restrict__ idx_out, float* __restrict__ lse_out, float* __restrict__ sel_key_out) {
    if (offset_dev) offset += *offset_dev;  // stream position kept on the device (CUDA-graph replay draws fresh noise)
    constexpr int kTrainChunk = train_chunk(kTrainWarps);
    __shared__ __align__(16) float wtab[kTrainChunk];     // exp(l - lmax)
    __shared__ __align__(16) float winv[kTrainChunk];     // exp(lmax - l)
    __shared__ float red[kTrainWarps];
    __shared__ float cand_v[kTrainWarps][kTrainCand];
    __shared__ int cand_i[kTrainWarps][kTrainCand];
    __shared__ int cand_n[kTrainWarps];
    __shared__ int win_i[kTrainWarps][8];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const float* logits_b = logits + (size_t)b * N;
    const int n_chunks = (N + kTrainChunk - 1) / kTrainChunk;

    // lmax and W = sum exp(l - lmax) over the pair
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, __ldg(logits_b + n));
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    DRB_UNROLL
    for (int w = 1; w < kTrainWarps; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float wsum = 0.f;
    auto build = [&](int c, bool accumulate) {
        const int base = c * kTrainChunk;
        const int len = min(kTrainChunk, N - base);
        const int len_pad = ((len + 127) / 128) * 128;
        for (int n = threadIdx.x; n < len_pad; n += blockDim.x) {
            const float d = n < len ? __ldg(logits_b + base + n) - mx : -INFINITY;
            const float w = __expf(d);
            wtab[n] = w;                         // 0 for the padding
            winv[n] = __expf(-d);                // +inf for the padding: log2(u) * inf = -inf never ranks
            if (accumulate) wsum += w;
        }
    };
    build(0, true);
    for (int n = kTrainChunk + threadIdx.x; n < N; n += blockDim.x) wsum += __expf(__ldg(logits_b + n) - mx);
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(FULL, wsum, o);
    if (lane == 0) {
        red[warp] = wsum;
        cand_n[warp] = 0;
    }
    __syncthreads();
    wsum = red[0];
    DRB_UNROLL
    for (int w = 1; w < kTrainWarps; ++w) wsum += red[w];
    const float x0 = -(3.f * S) / (0.6931471805599453f * wsum);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);

    // The CTA keeps its pair's tables and walks over groups of kTrainWarps hypotheses: the per-CTA set-up above (three
    // passes over the logits behind global-load latency and block barriers: 13 % of the instructions and a quarter of
    // the stall samples when every CTA served one group) is paid once per CTA, not once per group.
    for (int kg = blockIdx.x; kg * kTrainWarps < K; kg += gridDim.x) {
    const int k = kg * kTrainWarps + warp;
    const bool live = k < K;
    const long long row = (long long)b * K + k;
    if (kg != (int)blockIdx.x) {
        if (n_chunks > 1) {          // the sweep left the LAST chunk's tables behind
            __syncthreads();
            build(0, false);
            __syncthreads();
        }
        if (lane == 0) cand_n[warp] = 0;
        __syncwarp();
    }
    float top_v = -INFINITY, thr = -INFINITY, zacc = 0.f;
    int top_i = -1;
    // sweep 1: filtered, all chunks
    for (int c = 0; c < n_chunks; ++c) {
        if (c > 0) {
            __syncthreads();
            build(c, false);
            __syncthreads();
        }
        const int len = min(kTrainChunk, N - c * kTrainChunk);
        if (live)
            train_sweep<S, true>(wtab, winv, c * kTrainChunk, ((len + 127) / 128) * 128, (uint32_t)k, (uint32_t)b,
                                 (uint32_t)offset, k0, k1, x0, cand_v[warp], cand_i[warp], &cand_n[warp], zacc, top_v,
                                 top_i, thr, lane);
    }
    __syncwarp();
    const int nc = live ? cand_n[warp] : S;
    const bool redo = live && (nc < S || nc > kTrainCand);
    if (live && !redo) {
        // rank the listed candidates (value descending, index ascending on ties); lane holds entries lane, lane+32
        const float v0 = lane < nc ? cand_v[warp][lane] : -INFINITY, v1 = lane + 32 < nc ? cand_v[warp][lane + 32] : -INFINITY;
        const int i0 = lane < nc ? cand_i[warp][lane] : 0x7fffffff, i1 = lane + 32 < nc ? cand_i[warp][lane + 32] : 0x7fffffff;
        int r0 = 0, r1 = 0;
        for (int j = 0; j < nc; ++j) {
            const float vj = cand_v[warp][j];
            const int ij = cand_i[warp][j];
            r0 += (vj > v0 || (vj == v0 && ij < i0)) ? 1 : 0;
            r1 += (vj > v1 || (vj == v1 && ij < i1)) ? 1 : 0;
        }
        if (lane < nc && r0 < S) win_i[warp][r0] = i0;
        if (lane + 32 < nc && r1 < S) win_i[warp][r1] = i1;
        __syncwarp();
        if (lane < S) top_i = win_i[warp][lane];
    }
    // sweep 2 (rare): the hypotheses whose list came out short or overflowed, unfiltered.  With one chunk the tables
    // are still in place and the warp redoes its sweep on its own: no block barrier, so the warps of a CTA walk over
    // their groups independently (a barrier here makes every group as slow as its slowest warp).
    if (n_chunks == 1) {
        float zdummy = 0.f;
        if (redo)
            train_sweep<S, false>(wtab, winv, 0, ((N + 127) / 128) * 128, (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0,
                                  k1, x0, nullptr, nullptr, nullptr, zdummy, top_v, top_i, thr, lane);
    } else if (__syncthreads_or(redo ? 1 : 0)) {
        for (int c = 0; c < n_chunks; ++c) {
            if (n_chunks > 1) {
                __syncthreads();
                build(c, false);
                __syncthreads();
            }
            const int len = min(kTrainChunk, N - c * kTrainChunk);
            float zdummy = 0.f;
            if (redo)
                train_sweep<S, false>(wtab, winv, c * kTrainChunk, ((len + 127) / 128) * 128, (uint32_t)k, (uint32_t)b,
                                      (uint32_t)offset, k0, k1, x0, nullptr, nullptr, nullptr, zdummy, top_v, top_i, thr,
                                      lane);
        }
    }
    if (!live) continue;             // (a warp without a hypothesis still takes part in the block barriers above)
    DRB_UNROLL
    for (int o = 16; o > 0; o >>= 1) zacc += __shfl_xor_sync(FULL, zacc, o);
    // Z = sum w / e = -(1 / ln 2) * zacc ;  lse = lmax + ln Z
    if (lane == 0) lse_out[row] = mx + __logf(-1.4426950408889634f * zacc);
    int rank = 0;
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        const int oj = __shfl_sync(FULL, top_i, j);
        rank += (oj < top_i) ? 1 : 0;
    }
    if (lane < S) {
        idx_out[(size_t)row * S + rank] = top_i;
        // key of the selected point, from the same draw: l + g with g = -ln(e)
        const Philox4 r = philox4x32_10((uint32_t)(top_i >> 2), (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0, k1);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        uint32_t bits = rr[0];
        DRB_UNROLL
        for (int i = 1; i < 4; ++i) bits = ((top_i & 3) == i) ? rr[i] : bits;
        sel_key_out[(size_t)row * S + rank] = __ldg(logits_b + top_i) + gumbel_from_bits(bits);
    }
    }   // hypothesis groups of this CTA
}

template <int S>
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_kernel(const float* __restrict__ logits, const float* __restrict__ noise, uint64_t seed, uint64_t offset,
              const unsigned long long* __restrict__ offset_dev, float tau, int B, int K, int N,
              int32_t* __restrict__ idx_out, float* __restrict__ lse_out, float* __restrict__ sel_key_out,
              float* __restrict__ noise_out) {
    if (offset_dev) offset += *offset_dev;  // stream position kept on the device (CUDA-graph replay draws fresh noise)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kSamplerWarps + warp;  // b * K + k
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    const int k = (int)(row % K);
    const float* logits_b = logits + (size_t)b * N;
    const float* noise_row = noise ? noise + (size_t)row * N : nullptr;
    float* noise_out_row = noise_out ? noise_out + (size_t)row * N : nullptr;
    const bool vec = (N % 4 == 0);
    const unsigned FULL = 0xffffffffu;

    // running top-S, sorted descending, element j lives in lane j
    float top_v = -INFINITY;
    int top_i = -1;
    float thr = -INFINITY;
    // online log-sum-exp
    float run_m = -INFINITY, run_s = 0.f;

    for (int n0 = lane * 4; n0 < ((N + 127) / 128) * 128; n0 += 128) {
        KeyQuad q;
        if (n0 < N) {
            q = load_keys(logits_b, noise_row, noise_out_row, vec, n0, N, tau, (uint32_t)k, (uint32_t)b, seed, offset);
        } else {
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) q.k[i] = -INFINITY;
        }
        if (lse_out != nullptr) {
            const float m4 = fmaxf(fmaxf(q.k[0], q.k[1]), fmaxf(q.k[2], q.k[3]));
            if (m4 > -INFINITY) {
                const float nm = fmaxf(run_m, m4);
                float s = run_s * __expf(run_m - nm);
                DRB_UNROLL
                for (int i = 0; i < 4; ++i) s += __expf(q.k[i] - nm);
                run_m = nm;
                run_s = s;
            }
        }
        {
            float bv = q.k[0];
            int bi = 0;
            DRB_UNROLL
            for (int i = 1; i < 4; ++i)
                if (q.k[i] > bv) { bv = q.k[i]; bi = i; }
            merge_candidates<S>(bv, n0 + bi, top_v, top_i, thr, lane);
            float second = -INFINITY;
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) second = (i == bi) ? second : fmaxf(second, q.k[i]);
            if (__any_sync(FULL, second > thr)) {
                DRB_UNROLL
                for (int i = 0; i < 4; ++i)
                    merge_candidates<S>((i == bi) ? -INFINITY : q.k[i], n0 + i, top_v, top_i, thr, lane);
            }
        }
    }
    // ascending-index order of the boolean-mask gather (ransac.py:65)
    int rank = 0;
    DRB_UNROLL
    for (int j = 0; j < S; ++j) {
        const int oj = __shfl_sync(FULL, top_i, j);
        rank += (oj < top_i) ? 1 : 0;
    }
    if (lane < S) {
        idx_out[(size_t)row * S + rank] = top_i;
        if (sel_key_out) sel_key_out[(size_t)row * S + rank] = top_v;
    }
    if (lse_out != nullptr) {
        DRB_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(FULL, run_m, o);
            const float os = __shfl_xor_sync(FULL, run_s, o);
            const float nm = fmaxf(run_m, om);
            const float a = (run_m > -INFINITY) ? run_s * __expf(run_m - nm) : 0.f;
            const float c = (om > -INFINITY) ? os * __expf(om - nm) : 0.f;
            run_m = nm;
            run_s = a + c;
        }
        if (lane == 0) lse_out[row] = run_m + __logf(run_s);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward: dL/dlogits[n] += (1/tau) ( sum_{k: n selected} y[k,n] g[k,n]  -  sum_k y[k,n] c_k ),
// c_k = sum_j y[k, n_j] g_sel[k, j],  y[k,n] = exp(key[k,n] - lse[k]).
// Pass 1 (one thread per (b,k)): c_k and the sparse term (atomics on s entries).
// Pass 2 (one thread per 4 consecutive n, K split over blockIdx.y): the dense K x N term,
// regenerating / re-reading the noise; no K x N tensor is ever stored.
__global__ void sample_bwd_sparse_kernel(float tau, int B, int K, int N, int S, const int32_t* __restrict__ idx,
                                         const float* __restrict__ lse, const float* __restrict__ sel_key,
                                         const float* __restrict__ g_sel, float* __restrict__ ck,
                                         float* __restrict__ grad_logits) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * K) return;
    const int b = (int)(row / K);
    const float l = lse[row];
    float c = 0.f;
    const float it = 1.f / tau;
    for (int j = 0; j < S; ++j) {
        const float y = __expf(sel_key[row * S + j] - l);
        const float t = y * g_sel[row * S + j];
        c += t;
        atomicAdd(grad_logits + (size_t)b * N + idx[row * S + j], t * it);
    }
    ck[row] = c;
}

constexpr int kBwdThreads = 128;

__global__ void __launch_bounds__(kBwdThreads)
sample_bwd_dense_kernel(const float* __restrict__ logits, const float* __restrict__ noise, uint64_t seed,
                        uint64_t offset, const unsigned long long* __restrict__ offset_dev, float tau, int B, int K, int N,
                        int k_per_block, const float* __restrict__ lse, const float* __restrict__ ck,
                        float* __restrict__ grad_logits) {
    if (offset_dev) offset += *offset_dev;  // stream position kept on the device (CUDA-graph replay draws fresh noise)
    const int b = blockIdx.z;
    const int n0 = (blockIdx.x * kBwdThreads + threadIdx.x) * 4;
    const int k_begin = blockIdx.y * k_per_block;
    const int k_end = min(K, k_begin + k_per_block);
    if (n0 >= N) return;
    const float* logits_b = logits + (size_t)b * N;
    const bool vec = (N % 4 == 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise == nullptr && tau == 1.0f) {
        // Fast path (in-kernel noise, tau = 1):  y[k,n] = exp(l_n + g - lse_k) = exp(l_n - c) * exp(c - lse_k) / e,
        // e = -ln u the exponential behind the Gumbel draw; c = the largest of this thread's four logits, so the
        // per-element work is one log2 and one reciprocal instead of two logs and an exp.
        float l4[4], w4[4];
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) l4[i] = (n0 + i < N) ? __ldg(logits_b + n0 + i) : -INFINITY;
        const float c0 = fmaxf(fmaxf(l4[0], l4[1]), fmaxf(l4[2], l4[3]));
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) w4[i] = (n0 + i < N) ? __expf(l4[i] - c0) : 0.f;
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
        for (int k = k_begin; k < k_end; ++k) {
            const size_t row = (size_t)b * K + k;
            const Philox4 r = philox4x32_10((uint32_t)(n0 >> 2), (uint32_t)k, (uint32_t)b, (uint32_t)offset, k0, k1);
            // exp(c - lse_k) * c_k / ln 2, with the sign of 1 / log2(u) (< 0) folded in
            const float skc = -1.4426950408889634f * __expf(c0 - __ldg(lse + row)) * __ldg(ck + row);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            DRB_UNROLL
            for (int i = 0; i < 4; ++i) {
                const float lg = fminf(lg2_approx(uniform_from_bits(rr[i])), -1.4426950408889634e-10f);  // e >= 1e-10
                acc[i] = fmaf(w4[i] * rcp_fast(lg), skc, acc[i]);
            }
        }
    } else
    for (int k = k_begin; k < k_end; ++k) {
        const size_t row = (size_t)b * K + k;
        const float* noise_row = noise ? noise + row * N : nullptr;
        const KeyQuad q = load_keys(logits_b, noise_row, nullptr, vec, n0, N, tau, (uint32_t)k, (uint32_t)b, seed, offset);
        const float l = __ldg(lse + row);
        const float c = __ldg(ck + row);
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) acc[i] += __expf(q.k[i] - l) * c;
    }
    const float it = -1.f / tau;
    DRB_UNROLL
    for (int i = 0; i < 4; ++i)
        if (n0 + i < N) atomicAdd(grad_logits + (size_t)b * N + n0 + i, acc[i] * it);
}

}  // namespace drb

using namespace drb;

static int check_launch() { return cudaGetLastError() == cudaSuccess ? DRB_OK : DRB_ERR_CUDA; }

extern "C" int drb_sample(const float* logits, const float* noise, uint64_t seed, uint64_t offset,
                          const uint64_t* offset_dev, float tau, int B, int K, int N, int s, int32_t* idx, float* lse,
                          float* sel_key, float* noise_out, void* stream) {
    const unsigned long long* od = reinterpret_cast<const unsigned long long*>(offset_dev);
    if (!logits || !idx) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || s <= 0 || s > N || !(tau > 0.f)) return DRB_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * K;
    const unsigned grid = (unsigned)((rows + kSamplerWarps - 1) / kSamplerWarps);
    const dim3 block(kSamplerWarps * 32);
    if (!noise && lse && sel_key && !noise_out && tau == 1.0f && B <= 65535) {
        // training forward with in-kernel noise at tau = 1
        const bool one_chunk = N <= train_chunk(8);
        const int tw = one_chunk ? 8 : 16;
        // One group of `tw` hypotheses per CTA.  (Measured on the B200, round 2: capping the grid at one wave of CTAs
        // that walk over several groups with their tables built once is NOT faster -- 0.128 vs 0.124 ms at cfg5, 0.444
        // vs 0.432 at cfg3: with ~7 waves of short CTAs the set-up of one CTA hides under the sweeps of its
        // neighbours.  The kernel keeps the group loop; the launch does not use it.)
        const int groups = (K + tw - 1) / tw;
        const dim3 tgrid(groups, B);
#define DRB_LAUNCH_TRAIN(S_)                                                                                      \
    case S_:                                                                                                      \
        if (one_chunk)                                                                                            \
            sample_train_kernel<S_, 8><<<tgrid, 256, 0, st>>>(logits, seed, offset, od, K, N, idx, lse, sel_key); \
        else                                                                                                      \
            sample_train_kernel<S_, 16><<<tgrid, 512, 0, st>>>(logits, seed, offset, od, K, N, idx, lse, sel_key); \
        break;
        switch (s) {
            DRB_LAUNCH_TRAIN(3)
            DRB_LAUNCH_TRAIN(5)
            DRB_LAUNCH_TRAIN(7)
            DRB_LAUNCH_TRAIN(8)
            default:
                return DRB_ERR_UNSUPPORTED;
        }
#undef DRB_LAUNCH_TRAIN
        return check_launch();
    }
    if (!noise && !lse && !sel_key && !noise_out && N <= kRaceMaxN && B <= 65535) {
        // test mode: only the indices are wanted -> exponential-race fast path
        const dim3 rgrid((K + kSamplerWarps - 1) / kSamplerWarps, B);
#define DRB_LAUNCH_RACE(S_)                                                                   \
    case S_:                                                                                  \
        sample_race_kernel<S_><<<rgrid, block, 0, st>>>(logits, seed, offset, od, K, N, idx); \
        break;
        switch (s) {
            DRB_LAUNCH_RACE(3)
            DRB_LAUNCH_RACE(5)
            DRB_LAUNCH_RACE(7)
            DRB_LAUNCH_RACE(8)
            default:
                return DRB_ERR_UNSUPPORTED;
        }
#undef DRB_LAUNCH_RACE
        return check_launch();
    }
#define DRB_LAUNCH_SAMPLE(S_)                                                                                   \
    case S_:                                                                                                    \
        sample_kernel<S_><<<grid, block, 0, st>>>(logits, noise, seed, offset, od, tau, B, K, N, idx, lse,      \
                                                  sel_key, noise_out);                                          \
        break;
    switch (s) {
        DRB_LAUNCH_SAMPLE(3)
        DRB_LAUNCH_SAMPLE(5)
        DRB_LAUNCH_SAMPLE(7)
        DRB_LAUNCH_SAMPLE(8)
        default:
            return DRB_ERR_UNSUPPORTED;
    }
#undef DRB_LAUNCH_SAMPLE
    return check_launch();
}

extern "C" int drb_sample_sets(const float* logits, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int B,
                               int K, int N, int s, int32_t* idx, void* stream) {
    if (!logits || !idx) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || s <= 0 || s > N || B > 65535) return DRB_ERR_BAD_SHAPE;
    const size_t smem = (size_t)N * sizeof(float);
    if (smem > 200 * 1024) return DRB_ERR_UNSUPPORTED;  // the prefix sums of one pair must fit in shared memory
    const dim3 grid((K + kSetThreads - 1) / kSetThreads, B);
#define DRB_LAUNCH_SETS(S_)                                                                                       \
    case S_: {                                                                                                    \
        static std::atomic<unsigned long long> configured{0};                                                     \
        if (!ensure_dynamic_smem(sample_sets_kernel<S_>, 200 * 1024, configured)) return DRB_ERR_CUDA;            \
        sample_sets_kernel<S_><<<grid, kSetThreads, smem, (cudaStream_t)stream>>>(                                \
            logits, seed, offset, reinterpret_cast<const unsigned long long*>(offset_dev), K, N, idx);            \
        break;                                                                                                    \
    }
    switch (s) {
        DRB_LAUNCH_SETS(3)
        DRB_LAUNCH_SETS(5)
        DRB_LAUNCH_SETS(7)
        DRB_LAUNCH_SETS(8)
        default:
            return DRB_ERR_UNSUPPORTED;
    }
#undef DRB_LAUNCH_SETS
    return check_launch();
}

extern "C" int drb_sample_backward(const float* logits, const float* noise, uint64_t seed, uint64_t offset,
                                   const uint64_t* offset_dev, float tau, int B, int K, int N, int s,
                                   const int32_t* idx, const float* lse,
                                   const float* sel_key, const float* g_sel, float* scratch, float* grad_logits,
                                   void* stream) {
    if (!logits || !idx || !lse || !sel_key || !g_sel || !scratch || !grad_logits) return DRB_ERR_NULL_POINTER;
    if (B <= 0 || K <= 0 || N <= 0 || s <= 0 || !(tau > 0.f)) return DRB_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    float* ck = scratch;
    const long long rows = (long long)B * K;
    sample_bwd_sparse_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(tau, B, K, N, s, idx, lse, sel_key, g_sel,
                                                                             ck, grad_logits);
    const int n_blocks = (N + kBwdThreads * 4 - 1) / (kBwdThreads * 4);
    // spread K over enough blocks to fill the machine (148 SMs x several CTAs)
    int k_split = (148 * 8 + n_blocks * B - 1) / (n_blocks * B);
    if (k_split < 1) k_split = 1;
    if (k_split > K) k_split = K;
    const int k_per_block = (K + k_split - 1) / k_split;
    k_split = (K + k_per_block - 1) / k_per_block;
    dim3 grid(n_blocks, k_split, B);
    sample_bwd_dense_kernel<<<grid, kBwdThreads, 0, st>>>(logits, noise, seed, offset,
                                                          reinterpret_cast<const unsigned long long*>(offset_dev), tau,
                                                          B, K, N, k_per_block, lse, ck, grad_logits);
    return check_launch();
}
