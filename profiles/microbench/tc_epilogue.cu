// Microbenchmark (B200): the epilogue of the tensor-core scorer (csrc/score_tc.cu) with nothing around it -- no bulk
// copies, no MMAs, no mbarriers: EPI warps per CTA read a 128 x 256 accumulator from tensor memory chunk by chunk and
// run the pair-reciprocal arithmetic on it, tile after tile.  The real kernel's epilogue takes ~1440 cycles per tile
// with the MMAs switched off (profiles/r2_tc_ablate.jsonl), its instruction count says ~600; this tells which part of
// the difference is the epilogue's own (tcgen05.ld latency, dependent arithmetic, two warps per sub-partition) and
// which is the pipeline around it.  One CTA per SM; prints cycles per tile for every variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I differentiable_ransac_b200/csrc \
//        -o profiles/microbench/build/tc_epilogue profiles/microbench/tc_epilogue.cu && profiles/microbench/build/tc_epilogue
#include <cstdio>
#include <cuda_runtime.h>

#include "f32x2.cuh"
#include "tc_ptx.cuh"

using namespace drb;
using namespace drb::tc;

// MODE 0: tcgen05.ld + pair arithmetic (the kernel's)      1: tcgen05.ld only        2: arithmetic only (registers)
// MODE 3: folded arithmetic, FFMA.SAT + FADD2 (7 per step) 4: as 0 with x16 loads, two in flight
// HS 1: the kernel's accumulator handshake around every tile (wait d_full, fences, arrive d_empty), with the spare warp
// standing in for the MMA issuer (it waits for d_empty and arrives on d_full at once); HS 2: the same, the spare warp
// polling with nanosleep(100) between tries
template <int EPI, int MODE, int HS = 0>
__global__ void __launch_bounds__(32 * (EPI + 1), 1)
epilogue_kernel(float* __restrict__ out, long long* __restrict__ cycles, int tiles, float nci_in) {
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(8) uint64_t d_full[2], d_empty[2];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], EPI);
        }
        fence_barrier_init();
    }
    __shared__ float4 stand_in[8][32];          // MODE 2: a conflict-free shared-memory stand-in for the tensor-memory load
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == EPI) tmem_alloc(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (threadIdx.x < 256) {
        const int i = threadIdx.x >> 5, l = threadIdx.x & 31;
        stand_in[i][l] = make_float4(1e-3f * (float)(((l * 7 + i) & 31) - 16), 1e-3f * (float)(((l * 5 + i) & 31) - 15),
                                     0.4f + 0.001f * (float)((l + i) & 63), 0.5f + 0.001f * (float)((l + 3 * i) & 63));
    }
    if (warp < 4) {   // fill both accumulators with plausible (r0, r1, j1, j0) groups: r ~ 1e-3, j ~ 0.5
        for (int c = 0; c < 512; c += 16) {
            uint32_t v[16];
            for (int i = 0; i < 16; ++i) {
                const float x = (i & 2) ? 0.4f + 0.001f * (float)((lane + c + i) & 63) : 1e-3f * (float)(((lane * 7 + c + i) & 31) - 16);
                v[i] = __float_as_uint(x);
            }
            tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == EPI) {
        __syncthreads();
        if (HS && lane == 0) {
            Ring rd;
            for (int t = 0; t < tiles; ++t) {
                if (HS == 2) {
                    uint32_t done = 0;
                    while (true) {
                        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                     : "=r"(done) : "r"(smem_u32(&d_empty[rd.idx])), "r"(rd.phase ^ 1u) : "memory");
                        if (done) break;
                        __nanosleep(100);
                    }
                } else {
                    mbar_wait(&d_empty[rd.idx], rd.phase ^ 1u);
                }
                mbar_arrive(&d_full[rd.idx]);
                rd.advance(2);
            }
        }
        __syncthreads();
        tmem_dealloc(tmem_base, 512);
        return;
    }
    constexpr int kParts = EPI / 4, kCols = 256 / kParts, kChunks = kCols / 32, kAcc = kChunks * 8;
    const int quarter = warp & 3, part = warp >> 2;
    const float nci = nci_in;
    pk2 acc[kAcc];
    for (int i = 0; i < kAcc; ++i) acc[i] = pk2_splat(0.f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    Ring rdq;
    for (int t = 0; t < tiles; ++t) {
        if (HS) {
            mbar_wait(&d_full[rdq.idx], rdq.phase);
            __syncwarp();
            tc_fence_after();
        }
        const float one = (t + lane < 1 << 30) ? 1.f : 0.f;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((t & 1) * 256 + part * kCols);
        if (MODE == 4) {
            uint32_t va[16], vb[16];
            tmem_ld16(taddr, va);
#pragma unroll
            for (int h = 0; h < kCols / 16; ++h) {
                uint32_t* cur = (h & 1) ? vb : va;
                uint32_t* nxt = (h & 1) ? va : vb;
                tmem_ld_wait();
                if (h + 1 < kCols / 16) tmem_ld16(taddr + (uint32_t)((h + 1) * 16), nxt);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const pk2 R = pk2_make(__uint_as_float(cur[4 * q]), __uint_as_float(cur[4 * q + 1]));
                    const pk2 R2 = pk2_mul(R, R);
                    const float ja = __uint_as_float(cur[4 * q + 2]), jb = __uint_as_float(cur[4 * q + 3]);
                    const float tn = rcp_approx(ja * jb) * nci;
                    float w0, w1;
                    pk2_split(pk2_mul(R2, pk2_make(ja, jb)), w0, w1);
                    acc[h * 4 + q] = pk2_add(acc[h * 4 + q], pk2_make(fma_sat(w0, tn, one), fma_sat(w1, tn, one)));
                }
            }
            if (HS) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[rdq.idx]);
                rdq.advance(2);
            }
            continue;
        }
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            uint32_t v[32];
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 f;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                                 : "r"(smem_u32(&stand_in[i][lane])));
                    v[4 * i] = __float_as_uint(f.x); v[4 * i + 1] = __float_as_uint(f.y);
                    v[4 * i + 2] = __float_as_uint(f.z); v[4 * i + 3] = __float_as_uint(f.w);
                }
            } else {
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                tmem_ld_wait();
            }
            if (MODE == 1) {
                acc[c * 8] = pk2_add(acc[c * 8], pk2_make(__uint_as_float(v[0] ^ v[13]), __uint_as_float(v[31] ^ v[20])));
                continue;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const pk2 R = pk2_make(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
                const pk2 R2 = pk2_mul(R, R);
                const float ja = __uint_as_float(v[4 * q + 2]), jb = __uint_as_float(v[4 * q + 3]);
                float w0, w1;
                pk2_split(pk2_mul(R2, pk2_make(ja, jb)), w0, w1);
                if (MODE == 3) {
                    const float tn = rcp_approx(ja * jb);
                    acc[c * 8 + q] = pk2_add(acc[c * 8 + q], pk2_make(fma_sat(w0, tn, 1.f), fma_sat(w1, tn, 1.f)));
                } else {
                    const float tn = rcp_approx(ja * jb) * nci;
                    acc[c * 8 + q] = pk2_add(acc[c * 8 + q], pk2_make(fma_sat(w0, tn, one), fma_sat(w1, tn, one)));
                }
            }
        }
        if (HS) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[rdq.idx]);
            rdq.advance(2);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < kAcc; ++i) {
        float a, b;
        pk2_split(acc[i], a, b);
        s += a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int EPI, int MODE, int HS = 0>
static void run(const char* what, int tiles) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 32 * (EPI + 1));
    cudaMalloc(&cyc, sizeof(long long) * sms);
    for (int rep = 0; rep < 2; ++rep) epilogue_kernel<EPI, MODE, HS><<<sms, 32 * (EPI + 1)>>>(out, cyc, tiles, -5.0e5f);
    cudaError_t e = cudaDeviceSynchronize();
    long long* h = new long long[sms];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    long long mx = 0;
    for (int i = 0; i < sms; ++i) {
        mean += (double)h[i];
        mx = h[i] > mx ? h[i] : mx;
    }
    printf("{\"epi_warps\": %d, \"mode\": %d, \"handshake\": %d, \"what\": \"%s\", \"cycles_per_tile_mean\": %.1f, \"cycles_per_tile_max\": %.1f, \"err\": \"%s\"}\n",
           EPI, MODE, HS, what, mean / sms / tiles, (double)mx / tiles, cudaGetErrorString(e));
    fflush(stdout);
    delete[] h;
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    const int tiles = 4000;
    run<8, 0>("ld + pair arithmetic (the kernel's epilogue)", tiles);
    run<8, 1>("ld only", tiles);
    run<8, 2>("arithmetic only", tiles);
    run<8, 3>("ld + folded arithmetic (7 per step)", tiles);
    run<8, 4>("x16 loads, next load issued before the arithmetic", tiles);
    run<16, 0>("ld + pair arithmetic", tiles);
    run<16, 1>("ld only", tiles);
    run<16, 2>("arithmetic only", tiles);
    run<16, 3>("ld + folded arithmetic", tiles);
    run<16, 4>("x16 loads, two in flight", tiles);
    run<8, 0, 1>("ld + pair arithmetic + accumulator handshake (stand-in MMA warp polls)", tiles);
    run<8, 0, 2>("ld + pair arithmetic + accumulator handshake (stand-in MMA warp polls with nanosleep)", tiles);
    run<8, 1, 1>("ld only + handshake", tiles);
    run<16, 0, 1>("ld + pair arithmetic + handshake", tiles);
    run<4, 0>("ld + pair arithmetic, ONE warp per sub-partition", tiles);
    run<4, 2>("arithmetic only, ONE warp per sub-partition", tiles);
    return 0;
}
