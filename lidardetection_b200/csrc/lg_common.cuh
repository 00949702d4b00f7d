// lg_common.cuh -- status / error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(%zu B dynamic smem): %s", bytes, cudaGetErrorString(e));
        return (int)e;
    }
    return LG_OK;
}

}  // namespace lg
