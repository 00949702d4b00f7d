// peak_fp32.cu -- unrolled-FFMA microbenchmark: the measured FP32 (CUDA-core) roofline denominator for
// bench.py (MEASURED_PEAKS.json only carries HBM and bf16 tensor figures).  Not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) ffma_kernel(float* __restrict__ sink, const int iters, const float b, const float c) {
    float a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) a[k] = __fmaf_rn(a[k], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += a[k];
    if (s == 123456.789f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

// launches one kernel of `blocks` x 256 threads; flops = blocks*256*iters*16*CHAINS*2
extern "C" __attribute__((visibility("default"))) double peakfp32_launch(float* sink, int iters, int blocks, int chains,
                                                                         void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (chains == 8) ffma_kernel<8><<<blocks, 256, 0, st>>>(sink, iters, 0.99999f, 1e-6f);
    else ffma_kernel<4><<<blocks, 256, 0, st>>>(sink, iters, 0.99999f, 1e-6f);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return (double)blocks * 256.0 * (double)iters * 16.0 * (double)(chains == 8 ? 8 : 4) * 2.0;
}
