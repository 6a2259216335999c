// Microbenchmark (B200): what limits packed-fp32 FMA (FFMA2) throughput -- the FMA pipe (1 per 2 cycles
// per SM sub-partition) or register-file read bandwidth?  Every variant runs the same number of packed
// instructions per warp; only the operand pattern changes.  Prints cycles per warp-instruction per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rf ffma2_rf.cu && ./ffma2_rf
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pk2;
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
__device__ __forceinline__ pk2 mk(float lo, float hi) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

constexpr int NACC = 8;

template <int V>
__global__ void __launch_bounds__(128) k(const float* __restrict__ in, float* __restrict__ out, long long* cyc, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float s[12];
    for (int i = 0; i < 12; ++i) s[i] = in[(tid * 12 + i) & 4095];
    pk2 sp[12];
    for (int i = 0; i < 12; ++i) sp[i] = mk(s[i], s[i]);
    pk2 X[NACC], Y[NACC], D[NACC];
    float xs[NACC], ds[NACC];
    for (int i = 0; i < NACC; ++i) {
        X[i] = mk(in[(tid + i) & 4095], in[(tid + i + 7) & 4095]);
        Y[i] = mk(in[(tid + i + 50) & 4095], in[(tid + i + 57) & 4095]);
        D[i] = mk(in[(tid + i + 100) & 4095], in[(tid + i + 107) & 4095]);
        xs[i] = in[(tid + i) & 4095];
        ds[i] = in[(tid + i + 100) & 4095];
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (V == 0) {        // 3 distinct pairs, nothing shared with the previous instruction
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], X[i], Y[i], D[i]);
            } else if (V == 1) { // scalar-broadcast A (distinct per instruction), pair B, pair C
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[i], Y[i], D[i]);
            } else if (V == 2) { // same pair A in consecutive instructions (reuse candidate), pair B, pair C
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], X[0], Y[i], D[i]);
            } else if (V == 3) { // same scalar A in consecutive instructions, pair B, pair C
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[0], Y[i], D[i]);
            } else if (V == 4) { // scalar A, pair B, scalar C (the "t" form): D = s*Y + s'
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[i], D[i], sp[(i + 1) % 12]);
            } else if (V == 5) { // same scalars A and C across consecutive instructions
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[0], D[i], sp[1]);
            } else if (V == 6) { // packed multiply, two distinct pairs
#pragma unroll
                for (int i = 0; i < NACC; ++i) MUL2(D[i], X[i], D[i]);
            } else if (V == 7) { // square-accumulate: D = X*X + D
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], X[i], X[i], D[i]);
            } else if (V == 8) { // scalar FFMA, three distinct registers
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA1(ds[i], xs[i], s[i], ds[i]);
            } else if (V == 9) { // scalar FFMA, A shared by consecutive instructions
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA1(ds[i], s[0], xs[i], ds[i]);
            } else if (V == 10) { // same pair B (middle slot) shared, scalar A distinct
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[i], X[0], D[i]);
            } else if (V == 11) { // scalar A shared AND only two reads: D = s*D + s'
#pragma unroll
                for (int i = 0; i < NACC; ++i) FMA2(D[i], sp[0], D[i], Y[i]);
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(D[i]));
        acc += lo + hi + ds[i];
    }
    out[tid] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void spin(long long cycles, long long* out) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0) out[0] = clock64() - t0;
}

static double g_ghz = 0.0;

template <int V>
void run(const char* name, const float* in, float* out, long long* cyc, int warps_per_smsp) {
    int dev_sms;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<V>, 128, 0);
    if (occ < warps_per_smsp) {
        printf("{\"variant\": \"%s\", \"warps_per_smsp\": %d, \"skipped\": \"occupancy %d blocks/SM\"}\n", name, warps_per_smsp, occ);
        return;
    }
    const int blocks = dev_sms * warps_per_smsp;  // 128 threads = 4 warps = one per SMSP
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<V><<<blocks, 128>>>(in, out, cyc, 10);
    cudaEventRecord(e0);
    k<V><<<blocks, 128>>>(in, out, cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? (double)h[i] : mx;
    const double instr_per_smsp = (double)warps_per_smsp * iters * 4 * NACC;
    printf("{\"variant\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_instr_per_smsp_event\": %.3f, \"by_max_block_clock\": %.3f}\n",
           name, warps_per_smsp, ms * 1e-3 * g_ghz * 1e9 / instr_per_smsp, mx / instr_per_smsp);
    delete[] h;
}

int main() {
    float *in, *out;
    long long* cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&out, 148 * 16 * 128 * 4);
    cudaMalloc(&cyc, 148 * 16 * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 1.0f + 1e-6f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    {   // SM clock under load: spin for a known number of cycles, time with events
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        spin<<<148, 128>>>(20000000, cyc);
        cudaEventRecord(e0);
        spin<<<148, 128>>>(200000000, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        g_ghz = (double)c / (ms * 1e-3) / 1e9;
        printf("{\"sm_clock_ghz\": %.4f}\n", g_ghz);
    }
    for (int w : {2, 4, 8}) {
        run<0>("ffma2 pair,pair,pair distinct", in, out, cyc, w);
        run<1>("ffma2 scalar,pair,pair distinct", in, out, cyc, w);
        run<2>("ffma2 SAMEpair,pair,pair", in, out, cyc, w);
        run<3>("ffma2 SAMEscalar,pair,pair", in, out, cyc, w);
        run<4>("ffma2 scalar,pair,scalar distinct", in, out, cyc, w);
        run<5>("ffma2 SAMEscalar,pair,SAMEscalar", in, out, cyc, w);
        run<6>("fmul2 pair,pair", in, out, cyc, w);
        run<7>("ffma2 X,X,D (square-accumulate)", in, out, cyc, w);
        run<8>("ffma scalar 3 distinct", in, out, cyc, w);
        run<9>("ffma scalar SAME a", in, out, cyc, w);
        run<10>("ffma2 scalar,SAMEpair,pair", in, out, cyc, w);
        run<11>("ffma2 SAMEscalar,pairD,pair", in, out, cyc, w);
    }
    return 0;
}
