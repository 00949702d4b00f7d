// lg_common.cuh -- status / error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/lidargeom.h"

namespace lg {

// thread-local last-error text (lg_last_error_string); defined in lg_api.cu
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return LG_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size) instead of on every call: the attribute
// sticks, and the reference's per-frame call pattern makes the entry points latency-sensitive.  A cache of immutable facts,
// not state that matters; benign if two threads race (both set the same value).
inline bool smem_attr_cached(const void* fn, int dev, size_t bytes, bool store) {
    struct Slot {
        std::atomic<const void*> fn{nullptr};
        std::atomic<size_t> bytes{0};
    };
    static Slot table[8][64];
    if (dev < 0 || dev >= 8) return false;
    Slot* row = table[dev];
    const size_t h = (reinterpret_cast<uintptr_t>(fn) >> 4) % 64;
    for (int probe = 0; probe < 8; probe++) {
        Slot& sl = row[(h + probe) % 64];
        const void* cur = sl.fn.load(std::memory_order_acquire);
        if (cur == fn) {
            if (!store) return sl.bytes.load(std::memory_order_acquire) >= bytes;
            if (sl.bytes.load(std::memory_order_relaxed) < bytes) sl.bytes.store(bytes, std::memory_order_release);
            return true;
        }
        if (cur == nullptr) {
            if (!store) return false;
            const void* expect = nullptr;
            if (sl.fn.compare_exchange_strong(expect, fn, std::memory_order_acq_rel) || expect == fn) {
                if (sl.bytes.load(std::memory_order_relaxed) < bytes) sl.bytes.store(bytes, std::memory_order_release);
                return true;
            }
        }
    }
    return false;  // table full: fall back to setting the attribute every time
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kernel);
    if (smem_attr_cached(fn, dev, bytes, false)) return LG_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(%zu B dynamic smem): %s", bytes, cudaGetErrorString(e));
        return (int)e;
    }
#ifdef LG_CARVEOUT_MAX
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
#endif
    smem_attr_cached(fn, dev, bytes, true);
    return LG_OK;
}

}  // namespace lg
