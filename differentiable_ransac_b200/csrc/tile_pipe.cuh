// Two-stage shared-memory pipeline for streaming a pair's correspondences past a
// CTA of model-owning threads: tiles are fetched with the TMA bulk-copy engine
// (cp.async.bulk, SASS UBLKCP) signalled through an mbarrier; a plain cooperative
// copy is used for a tile whose byte count or source address is not 16-byte
// aligned (odd tails of 24-byte 3-D correspondences).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace drb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Streams `n_items` items of ITEM_FLOATS floats each from `src` through two smem
// stages of TILE items.  Usage (all threads of the CTA):
//   TilePipe<...> pipe(smem_tiles, bars, src, n_items);  pipe.prologue();
//   for (t = 0; t < pipe.n_tiles; ++t) { const float* tile = pipe.acquire(t); ... ; pipe.release(t); }
template <int ITEM_FLOATS, int TILE>
struct TilePipe {
    float* base;    // two stages of TILE items each
    uint64_t* bar;  // two mbarriers
    const float* src;
    int n_items;
    int n_tiles;

    __device__ __forceinline__ TilePipe(float* smem_tiles, uint64_t* bars, const float* src_, int n_items_)
        : base(smem_tiles), bar(bars), src(src_), n_items(n_items_) {
        n_tiles = (n_items + TILE - 1) / TILE;
    }
    __device__ __forceinline__ float* stage(int t) const { return base + (t & 1) * (TILE * ITEM_FLOATS); }
    __device__ __forceinline__ int tile_items(int t) const {
        const int rem = n_items - t * TILE;
        return rem < TILE ? rem : TILE;
    }
    __device__ __forceinline__ bool bulk_ok(int t) const {
        const uint32_t bytes = (uint32_t)tile_items(t) * ITEM_FLOATS * 4u;
        const uintptr_t a = reinterpret_cast<uintptr_t>(src + (size_t)t * TILE * ITEM_FLOATS);
        return (bytes % 16u == 0u) && (a % 16u == 0u);
    }
    __device__ __forceinline__ void issue(int t) {  // called by thread 0 only
        if (t < n_tiles && bulk_ok(t)) {
            const uint32_t bytes = (uint32_t)tile_items(t) * ITEM_FLOATS * 4u;
            mbar_expect_tx(&bar[t & 1], bytes);
            bulk_g2s(stage(t), src + (size_t)t * TILE * ITEM_FLOATS, bytes, &bar[t & 1]);
        }
    }
    __device__ __forceinline__ void prologue() {
        if (threadIdx.x == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            issue(0);
            issue(1);
        }
    }
    __device__ __forceinline__ const float* acquire(int t) {
        if (bulk_ok(t)) {
            mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1));
        } else {
            const int nf = tile_items(t) * ITEM_FLOATS;
            const float* s = src + (size_t)t * TILE * ITEM_FLOATS;
            for (int i = threadIdx.x; i < nf; i += blockDim.x) stage(t)[i] = __ldg(s + i);
            __syncthreads();
        }
        return stage(t);
    }
    __device__ __forceinline__ void release(int t) {
        __syncthreads();  // every thread is done reading stage t & 1
        if (threadIdx.x == 0) {
            fence_proxy_async();
            issue(t + 2);
        }
    }
};

}  // namespace drb
