// emu_geom.cpp -- TEST INFRASTRUCTURE: host build of the device per-pair code (see lg_host_emu.h).
#include "lg_host_emu.h"
#include "../../lidardetection_b200/csrc/lg_geom.cuh"

using namespace lg;

template <int FL>
static void run(const float* a, int64_t n, const float* b, int64_t m, float* out, int mode) {
    float4* ra = new float4[(size_t)n * REC_F4];
    float4* rb = new float4[(size_t)m * REC_F4];
    for (int64_t i = 0; i < n; i++) make_record<FL>(a + i * 7, ra + i * REC_F4);
    for (int64_t j = 0; j < m; j++) make_record<FL>(b + j * 7, rb + j * REC_F4);
    float2 slab[16];
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            const float4* A = ra + i * REC_F4;
            const float4* B = rb + j * REC_F4;
            float v;
            if (mode & 4) {  // with the exact-zero cull in front, as the kernels use it
                v = cull_survives(A[2], B[2]) ? overlap_area<FL>(A, B, slab, 1) : 0.f;
            } else {
                v = overlap_area<FL>(A, B, slab, 1);
            }
            if ((mode & 3) == 1) v = iou_from_overlap(v, A[2].w, B[2].w);
            if ((mode & 3) == 2) v = iou3d_from_overlap(v, A[4], B[4]);
            out[i * m + j] = v;
        }
    delete[] ra;
    delete[] rb;
}

// mode: 0 overlap, 1 iou_bev, 2 iou3d; +4 = apply the circle cull first
extern "C" void emu_pairs(const float* a, int64_t n, const float* b, int64_t m, float* out, int mode, int flavor) {
    if (flavor) run<1>(a, n, b, m, out, mode);
    else run<0>(a, n, b, m, out, mode);
}
