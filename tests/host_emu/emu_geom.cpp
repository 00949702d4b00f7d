// emu_geom.cpp -- TEST INFRASTRUCTURE: host build of the device per-pair code (see lg_host_emu.h).
#include "lg_host_emu.h"
#include "../../lidardetection_b200/csrc/lg_geom.cuh"

using namespace lg;

static long long g_slow = 0;
extern "C" long long emu_slow_count() { return g_slow; }

template <int FL>
static void run(const float* a, int64_t n, const float* b, int64_t m, float* out, int mode) {
    float4* ra = new float4[(size_t)n * REC_F4];
    float4* rb = new float4[(size_t)m * REC_F4];
    for (int64_t i = 0; i < n; i++) make_record<FL>(a + i * 7, ra + i * REC_F4);
    for (int64_t j = 0; j < m; j++) make_record<FL>(b + j * 7, rb + j * REC_F4);
    float2 slab[16];
    float ang[16];
    auto slab16 = [&](int k) -> float2& { return slab[k]; };
    auto ang16 = [&](int k) -> float& { return ang[k]; };
    auto pair_area = [&](const float4* A, const float4* B) {
        float v = overlap_area<FL>(A, B, slab, 1, 1u);
        if (v < 0.f) {
            g_slow++;
            v = overlap_area_slow<FL>(A, B, slab16, ang16);
        }
        return v;
    };
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            const float4* A = ra + i * REC_F4;
            const float4* B = rb + j * REC_F4;
            float v;
            if (mode & 4) {  // with the exact-zero cull in front, as the kernels use it
                v = cull_survives(A[REC_CULL], B[REC_CULL]) ? pair_area(A, B) : 0.f;
            } else {
                v = pair_area(A, B);
            }
            if ((mode & 3) == 1) v = iou_from_overlap(v, A[REC_CULL].w, B[REC_CULL].w);
            if ((mode & 3) == 2) v = iou3d_from_overlap(v, A[REC_Z], B[REC_Z]);
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

// diagnostics: vertices (insertion order), centroid and pseudo-angle keys of one pair (CUDA flavor)
extern "C" int emu_debug_pair(const float* a, const float* b, float* verts /*[16][2]*/, float* centroid /*[2]*/, uint32_t* keys /*[16]*/) {
    float4 A[REC_F4], B[REC_F4];
    make_record<1>(a, A);
    make_record<1>(b, B);
    uint32_t xmask, cmask;
    pair_masks<1>(A, B, xmask, cmask);
    int n = 0;
    float sx = 0.f, sy = 0.f;
    while (xmask) {
        const int e = __ffs(xmask) - 1;
        xmask &= xmask - 1;
        const float2 v = crossing_point<1>(A, B, e >> 2, e & 3);
        if (n < 16) { verts[2 * n] = v.x; verts[2 * n + 1] = v.y; }
        sx += v.x; sy += v.y; n++;
    }
    while (cmask) {
        const int e = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float4 c = (e & 1) ? A[e >> 1] : B[e >> 1];
        if (n < 16) { verts[2 * n] = c.x; verts[2 * n + 1] = c.y; }
        sx += c.x; sy += c.y; n++;
    }
    const float inv = 1.0f / (float)n;
    centroid[0] = sx * inv; centroid[1] = sy * inv;
    for (int k = 0; k < n && k < 16; k++) keys[k] = angle_key(verts[2 * k], verts[2 * k + 1], centroid[0], centroid[1], k);
    return n;
}
