// Per-DEVICE launch configuration (host side).  cudaFuncSetAttribute and the SM count belong to the current
// device, and a process may drive several (one rank per GPU is the norm, but nothing forbids one process touching
// two): every cache here is indexed by the current device and is safe to race on (the worst case is a repeated,
// idempotent runtime call).
#pragma once

#include <cuda_runtime.h>

#include <atomic>

namespace drb {

constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev % kMaxDevices;
}

// Opt a kernel into `bytes` of dynamic shared memory on the current device, once per (call site, device).
// `done` is the call site's own static bit mask.
template <class Kernel>
inline bool ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<unsigned long long>& done) {
    const unsigned long long bit = 1ull << current_device();
    if (done.load(std::memory_order_acquire) & bit) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    done.fetch_or(bit, std::memory_order_release);
    return true;
}

// SM count of the current device (148 on a B200), cached per device.
inline int sm_count_current_device() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    n = 148;
    int real = 0;
    cudaGetDevice(&real);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, real);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

}  // namespace drb
