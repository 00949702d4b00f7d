// lg_geom.cuh -- per-box records and the per-pair rotated-rectangle overlap for sm_100a.
//
// Behavioural spec: the reference's box_overlap / iou_bev
//   /root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:15-234  (CUDA build)
//   /root/reference/pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:39-229        (CPU build)
// This is not a translation of that code.  What is kept is the *arithmetic contract* -- the exact
// FP32 expression (and, for FL = 1, the FMA contraction ptxas applies to the reference kernels on
// sm_100a; DESIGN.md lists every decoded site) of each value that reaches the result: corner
// coordinates, the four straddle determinants, the crossing point, the margin test, the fan area,
// the IoU quotient.  Everything else is re-designed for the B200:
//   * all per-box work (4 sincos, corners, margin-expanded half extents, areas, z-range, volume,
//     a conservative cull radius) is hoisted into an 80-byte record computed once per box instead
//     of once per pair (the reference re-evaluates 20 sinf/cosf per pair);
//   * polygon vertices live in a per-thread shared-memory slab, not in local memory;
//   * the 16 straddle tests and 8 margin tests run uniformly into bit masks; only accepted crossings
//     are expanded, in a compact loop, so warps do not serialise over 16 divergent branch bodies;
//   * the angular order comes from a monotone pseudo-angle packed with the vertex index into one
//     32-bit key and an 8- (or, rarely, 16-) input min/max sorting network (the reference: ~24 atan2f
//     + bubble sort).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// full-precision libdevice sinf/cosf (no -use_fast_math); the CPU-only test tier re-points these at the
// oracle's restatement of libdevice when it compiles this header for the host (tests/host_emu/)
#ifndef LG_SINF
#define LG_SINF(x) sinf(x)
#define LG_COSF(x) cosf(x)
#endif

namespace lg {

// ---- record layout (20 floats = 5 x float4 = 80 B) -------------------------------------------
//  [0..7]   rotated corners x0,y0,x1,y1,x2,y2,x3,y3   order (-,-),(+,-),(+,+),(-,+)
//  [8,9]    centre x, y
//  [10]     cull radius (conservative)                -- rec[2] alone feeds the cull phase
//  [11]     dx*dy
//  [12,13]  cos(-heading), sin(-heading)              (check_in_box2d recomputes trig of the negated angle)
//  [14,15]  dx/2 + 1e-2, dy/2 + 1e-2                  (MARGIN = 1e-2, kernel.cu:53)
//  [16,17]  z + dz/2, z - dz/2                        (iou3d_nms_utils.py:60-63)
//  [18]     dx*dy*dz
//  [19]     unused
constexpr int REC_FLOATS = 20;
constexpr int REC_F4 = 5;

// ---- arithmetic contract helpers --------------------------------------------------------------
// FL = 1: reference CUDA build (nvcc 12.9, sm_100a):  a*b - c*d  ==  fma(a, b, -(c*d))
// FL = 0: reference CPU build (g++ -O2, x86-64):      every operation individually rounded
// __fmul_rn/__fmaf_rn are never re-contracted by nvcc/ptxas, so both are exactly what is written.
template <int FL>
__device__ __forceinline__ float msub(float a, float b, float c, float d) {  // a*b - c*d
    if (FL) return __fmaf_rn(a, b, -__fmul_rn(c, d));
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
template <int FL>
__device__ __forceinline__ float madd_first(float a, float b, float c, float d) {  // a*b + c*d, first fused
    if (FL) return __fmaf_rn(a, b, __fmul_rn(c, d));
    return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
template <int FL>
__device__ __forceinline__ float madd_second(float a, float b, float c, float d) {  // a*b + c*d, second fused
    if (FL) return __fmaf_rn(c, d, __fmul_rn(a, b));
    return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}

// ---- per-box record ----------------------------------------------------------------------------
template <int FL>
__device__ __forceinline__ void make_record(const float* __restrict__ box, float4* __restrict__ rec) {
    const float cx = box[0], cy = box[1], z = box[2], dx = box[3], dy = box[4], dz = box[5], th = box[6];
    const float hx = __fmul_rn(dx, 0.5f), hy = __fmul_rn(dy, 0.5f);
    const float x1 = __fsub_rn(cx, hx), x2 = __fadd_rn(cx, hx), y1 = __fsub_rn(cy, hy), y2 = __fadd_rn(cy, hy);
    const float co = LG_COSF(th), si = LG_SINF(th);
    // rotate_around_center works on (corner - centre), which is NOT exactly +-h after rounding
    const float ex1 = __fsub_rn(x1, cx), ex2 = __fsub_rn(x2, cx), ey1 = __fsub_rn(y1, cy), ey2 = __fsub_rn(y2, cy);
    float X[4], Y[4];
    const float ddx[4] = {ex1, ex2, ex2, ex1};
    const float ddy[4] = {ey1, ey1, ey2, ey2};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (FL) {
            X[k] = __fadd_rn(cx, __fmaf_rn(co, ddx[k], -__fmul_rn(si, ddy[k])));
            Y[k] = __fadd_rn(cy, __fmaf_rn(si, ddx[k], __fmul_rn(co, ddy[k])));
        } else {
            X[k] = __fadd_rn(__fadd_rn(__fmul_rn(ddx[k], co), __fmul_rn(ddy[k], -si)), cx);
            Y[k] = __fadd_rn(__fadd_rn(__fmul_rn(ddx[k], si), __fmul_rn(ddy[k], co)), cy);
        }
    }
    const float cn = LG_COSF(-th), sn = LG_SINF(-th);
    const float tx = __fadd_rn(hx, 1e-2f), ty = __fadd_rn(hy, 1e-2f);
    const float area = __fmul_rn(dx, dy);
    // Conservative exact-zero cull.  A pair yields a polygon vertex only if an edge of
    // A crosses an edge of B, or a corner of one box lies in the other's margin-expanded rectangle; both
    // need |ca - cb| <= ra + rb with r = sqrt(tx^2 + ty^2) (tx,ty = half extents + 1 cm).  Slack on top:
    // 0.01 % + 1 mm + 1e-6 * |centre|, orders of magnitude above the few-ulp error of the rotated corners.
    const float rad = sqrtf(tx * tx + ty * ty) * 1.0001f + 1e-3f + 1e-6f * (fabsf(cx) + fabsf(cy));
    rec[0] = make_float4(X[0], Y[0], X[1], Y[1]);
    rec[1] = make_float4(X[2], Y[2], X[3], Y[3]);
    rec[2] = make_float4(cx, cy, rad, area);
    rec[3] = make_float4(cn, sn, tx, ty);
    rec[4] = make_float4(__fadd_rn(z, __fmul_rn(dz, 0.5f)), __fsub_rn(z, __fmul_rn(dz, 0.5f)), __fmul_rn(area, dz), 0.f);
}

// exact-zero cull on the rec[2] quads (cx, cy, rad, area): false => the reference returns exactly +0.0
__device__ __forceinline__ bool cull_survives(const float4 ac, const float4 bc) {
    const float dx = ac.x - bc.x, dy = ac.y - bc.y, rr = ac.z + bc.z;
    return !(dx * dx + dy * dy > rr * rr);  // NaN => keep: the polygon path decides
}

// ---- sorting networks on packed 32-bit keys (verified with the 0-1 principle, tools/check_networks.py) ----
__device__ __forceinline__ void cex(uint32_t& a, uint32_t& b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}
__device__ __forceinline__ void sort8(uint32_t (&k)[8]) {  // 19 compare-exchanges
    cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]);
    cex(k[0], k[4]); cex(k[1], k[5]); cex(k[2], k[6]); cex(k[3], k[7]);
    cex(k[0], k[1]); cex(k[2], k[3]); cex(k[4], k[5]); cex(k[6], k[7]);
    cex(k[2], k[4]); cex(k[3], k[5]);
    cex(k[1], k[4]); cex(k[3], k[6]);
    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);
}
__device__ __forceinline__ void sort16(uint32_t (&k)[16]) {  // Batcher odd-even merge sort, 63 compare-exchanges
    cex(k[0], k[1]); cex(k[2], k[3]); cex(k[0], k[2]); cex(k[1], k[3]); cex(k[1], k[2]); cex(k[4], k[5]);
    cex(k[6], k[7]); cex(k[4], k[6]); cex(k[5], k[7]); cex(k[5], k[6]); cex(k[0], k[4]); cex(k[2], k[6]);
    cex(k[2], k[4]); cex(k[1], k[5]); cex(k[3], k[7]); cex(k[3], k[5]); cex(k[1], k[2]); cex(k[3], k[4]);
    cex(k[5], k[6]); cex(k[8], k[9]); cex(k[10], k[11]); cex(k[8], k[10]); cex(k[9], k[11]); cex(k[9], k[10]);
    cex(k[12], k[13]); cex(k[14], k[15]); cex(k[12], k[14]); cex(k[13], k[15]); cex(k[13], k[14]); cex(k[8], k[12]);
    cex(k[10], k[14]); cex(k[10], k[12]); cex(k[9], k[13]); cex(k[11], k[15]); cex(k[11], k[13]);
    cex(k[9], k[10]); cex(k[11], k[12]); cex(k[13], k[14]); cex(k[0], k[8]); cex(k[4], k[12]); cex(k[4], k[8]);
    cex(k[2], k[10]); cex(k[6], k[14]); cex(k[6], k[10]); cex(k[2], k[4]); cex(k[6], k[8]); cex(k[10], k[12]);
    cex(k[1], k[9]); cex(k[5], k[13]); cex(k[5], k[9]); cex(k[3], k[11]); cex(k[7], k[15]); cex(k[7], k[11]);
    cex(k[3], k[5]); cex(k[7], k[9]); cex(k[11], k[13]); cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);
    cex(k[7], k[8]); cex(k[9], k[10]); cex(k[11], k[12]); cex(k[13], k[14]);
}

// Monotone stand-in for atan2f(dy, dx) on (-pi, pi]: copysign(1 - dx/(|dx|+|dy|), dy) in [-2, 2],
// quantised to 2^-25 and packed above the 4-bit vertex index (ties keep insertion order, like the
// reference's stable bubble sort).  Only the ORDER of the vertices depends on it.
__device__ __forceinline__ uint32_t angle_key(float px, float py, float cx, float cy, int idx) {
    const float dx = px - cx, dy = py - cy;
    const float ad = fabsf(dx) + fabsf(dy);
    float k = 1.0f - __fdividef(dx, ad);
    k = (ad > 0.f) ? k : 0.f;  // atan2f(0, 0) = 0
    k = copysignf(k, dy);
    const uint32_t q = __float2uint_rn(__fmaf_rn(k, 33554432.0f, 67108864.0f));  // (k + 2) * 2^25 <= 2^27
    return (q << 4) | (uint32_t)idx;
}

// fan area from the first vertex in angular order (kernel.cu:219-224): NK = 8 or 16 sorted keys
template <int FL, int NK>
__device__ __forceinline__ float fan_area(const uint32_t (&key)[NK], const int cnt, const float2* __restrict__ slab,
                                          const int sstride) {
    const float2 p0 = slab[(key[0] & 15) * sstride];
    const float2 p1 = slab[(key[1] & 15) * sstride];
    float ux = p1.x - p0.x, uy = p1.y - p0.y;
    float area = 0.f;  // term 0 is cross(0, u) == +-0 exactly
#pragma unroll
    for (int k = 1; k < NK - 1; k++) {
        if (k + 1 < cnt) {
            const float2 pn = slab[(key[k + 1] & 15) * sstride];
            const float vx = pn.x - p0.x, vy = pn.y - p0.y;
            area = __fadd_rn(area, msub<FL>(ux, vy, uy, vx));
            ux = vx;
            uy = vy;
        }
    }
    return area;
}

// ---- the pair -----------------------------------------------------------------------------------
// A, B: records in SHARED memory (16-byte aligned; corners are re-read with dynamic indices).
// slab: this thread's vertex column, entry k at slab[k * sstride].
// Returns the overlap area exactly as the reference defines it (kernel.cu:104-225).
//
// Structure (all lanes of a warp stay converged except in the two short, compact loops):
//   A. uniform: the 16 edge x edge straddle tests -> a 16-bit mask of accepted crossings; the 8 corner
//      margin tests -> an 8-bit mask;
//   B. loop over the set crossing bits (ascending = the reference's i-outer / j-inner order): recompute
//      the determinants of that one edge pair and emit the crossing point;
//   C. append the flagged corners (reference order: B corner k, then A corner k);
//   D. pseudo-angle keys, 8- or 16-input sorting network, fan area.
template <int FL>
__device__ float overlap_area(const float4* __restrict__ A, const float4* __restrict__ B, float2* __restrict__ slab,
                              const int sstride) {
    float ax[4], ay[4], bx[4], by[4];
    {
        const float4 a0 = A[0], a1 = A[1], b0 = B[0], b1 = B[1];
        ax[0] = a0.x; ay[0] = a0.y; ax[1] = a0.z; ay[1] = a0.w; ax[2] = a1.x; ay[2] = a1.y; ax[3] = a1.z; ay[3] = a1.w;
        bx[0] = b0.x; by[0] = b0.y; bx[1] = b0.z; by[1] = b0.w; bx[2] = b1.x; by[2] = b1.y; bx[3] = b1.z; by[3] = b1.w;
    }
    // ---- A. straddle tests (kernel.cu:43-49, 67-75)
    uint32_t xmask = 0;
    {
        float eminx[4], emaxx[4], eminy[4], emaxy[4], fx[4], fy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int j1 = (j + 1) & 3;
            eminx[j] = fminf(bx[j], bx[j1]); emaxx[j] = fmaxf(bx[j], bx[j1]);
            eminy[j] = fminf(by[j], by[j1]); emaxy[j] = fmaxf(by[j], by[j1]);
            fx[j] = bx[j1] - bx[j]; fy[j] = by[j1] - by[j];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int i1 = (i + 1) & 3;
            const float p0x = ax[i], p0y = ay[i], p1x = ax[i1], p1y = ay[i1];
            const float pminx = fminf(p0x, p1x), pmaxx = fmaxf(p0x, p1x), pminy = fminf(p0y, p1y), pmaxy = fmaxf(p0y, p1y);
            const float ex = p1x - p0x, ey = p1y - p0y;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int j1 = (j + 1) & 3;
                const float q0x = bx[j], q0y = by[j], q1x = bx[j1], q1y = by[j1];
                const bool rc = pminx <= emaxx[j] && eminx[j] <= pmaxx && pminy <= emaxy[j] && eminy[j] <= pmaxy;
                const float s1 = msub<FL>(q0x - p0x, ey, ex, q0y - p0y);                         // cross(q0, p1, p0)
                const float s2 = __fsub_rn(__fmul_rn(ex, q1y - p0y), __fmul_rn(q1x - p0x, ey));  // cross(p1, q1, p0)
                const float s3 = msub<FL>(p0x - q0x, fy[j], fx[j], p0y - q0y);                   // cross(p0, q1, q0)
                const float s4 = msub<FL>(fx[j], p1y - q0y, p1x - q0x, fy[j]);                   // cross(q1, p1, q0)
                if (rc && __fmul_rn(s1, s2) > 0.f && __fmul_rn(s3, s4) > 0.f) xmask |= 1u << (i * 4 + j);
            }
        }
    }
    // ---- corner margin tests (kernel.cu:51-61): bit 2k = B corner k in A, bit 2k+1 = A corner k in B
    uint32_t cmask = 0;
    {
        const float4 am = A[2], at = A[3], bm = B[2], bt = B[3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            {
                const float dx = bx[k] - am.x, dy = by[k] - am.y;
                const float rx = msub<FL>(dx, at.x, dy, at.y);
                const float ry = madd_second<FL>(dx, at.y, dy, at.x);
                if (fabsf(rx) < at.z && fabsf(ry) < at.w) cmask |= 1u << (2 * k);
            }
            {
                const float dx = ax[k] - bm.x, dy = ay[k] - bm.y;
                const float rx = msub<FL>(dx, bt.x, dy, bt.y);
                const float ry = madd_second<FL>(dx, bt.y, dy, bt.x);
                if (fabsf(rx) < bt.z && fabsf(ry) < bt.w) cmask |= 1u << (2 * k + 1);
            }
        }
    }
    const int cnt = __popc(xmask) + __popc(cmask);  // <= 16 + 8; geometrically <= 16
    if (cnt <= 2) return 0.f;                        // the fan sum is empty or a single zero term

    // ---- B. crossing points (kernel.cu:77-91), ascending bit order == reference append order
    const float2* __restrict__ Ac = reinterpret_cast<const float2*>(A);
    const float2* __restrict__ Bc = reinterpret_cast<const float2*>(B);
    int n = 0;
    float sx = 0.f, sy = 0.f;
    while (xmask) {
        const int e = __ffs(xmask) - 1;
        xmask &= xmask - 1;
        const int i = e >> 2, j = e & 3;
        const float2 p0 = Ac[i], p1 = Ac[(i + 1) & 3], q0 = Bc[j], q1 = Bc[(j + 1) & 3];
        const float ex = p1.x - p0.x, ey = p1.y - p0.y;
        const float s1 = msub<FL>(q0.x - p0.x, ey, ex, q0.y - p0.y);
        const float t72 = __fmul_rn(ex, q1.y - p0.y), t73 = __fmul_rn(q1.x - p0.x, ey);
        const float s5 = __fsub_rn(t73, t72);  // cross(q1, p1, p0): its two products are shared with s2, unfused
        const float den = __fsub_rn(s5, s1);
        float X, Y;
        if (fabsf(den) > 1e-8f) {
            X = __fdiv_rn(msub<FL>(s5, q0.x, s1, q1.x), den);
            Y = __fdiv_rn(msub<FL>(s5, q0.y, s1, q1.y), den);
        } else {
            const float a0 = p0.y - p1.y, b0 = ex, c0 = msub<FL>(p0.x, p1.y, p1.x, p0.y);
            const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = msub<FL>(q0.x, q1.y, q1.x, q0.y);
            const float D = msub<FL>(a0, b1, a1, b0);
            X = __fdiv_rn(msub<FL>(b0, c1, b1, c0), D);
            Y = __fdiv_rn(msub<FL>(a1, c0, a0, c1), D);
        }
        if (n < 16) slab[n * sstride] = make_float2(X, Y);
        sx += X;
        sy += Y;
        n++;
    }
    // ---- C. corners
    while (cmask) {
        const int e = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float2 c = (e & 1) ? Ac[e >> 1] : Bc[e >> 1];
        if (n < 16) slab[n * sstride] = c;
        sx += c.x;
        sy += c.y;
        n++;
    }
    const int m = min(cnt, 16);
    const float inv = __fdividef(1.0f, (float)cnt);
    const float mx = sx * inv, my = sy * inv;  // centroid: only orders the vertices

    // ---- D. order + fan
    float area;
    if (m <= 8) {
        uint32_t key[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            key[k] = 0xFFFFFFF0u | k;
            if (k < m) {
                const float2 p = slab[k * sstride];
                key[k] = angle_key(p.x, p.y, mx, my, k);
            }
        }
        sort8(key);
        area = fan_area<FL, 8>(key, m, slab, sstride);
    } else {  // near-coincident boxes (~0.5 % of overlapping pairs)
        uint32_t key[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            key[k] = 0xFFFFFFF0u | k;
            if (k < m) {
                const float2 p = slab[k * sstride];
                key[k] = angle_key(p.x, p.y, mx, my, k);
            }
        }
        sort16(key);
        area = fan_area<FL, 16>(key, m, slab, sstride);
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// iou_bev (kernel.cu:227-234)
__device__ __forceinline__ float iou_from_overlap(float s, float sa, float sb) {
    return __fdiv_rn(s, fmaxf(__fsub_rn(__fadd_rn(sa, sb), s), 1e-8f));
}

// boxes_iou3d_gpu's epilogue (iou3d_nms_utils.py:69-79), each torch op individually rounded
__device__ __forceinline__ float iou3d_from_overlap(float ov, const float4 az, const float4 bz) {
    const float max_of_min = fmaxf(az.y, bz.y), min_of_max = fminf(az.x, bz.x);
    const float h = fmaxf(__fsub_rn(min_of_max, max_of_min), 0.f);
    const float o3d = __fmul_rn(ov, h);
    const float den = fmaxf(__fsub_rn(__fadd_rn(az.z, bz.z), o3d), 1e-6f);
    return __fdiv_rn(o3d, den);
}

// axis-aligned BEV IoU of nms_normal (kernel.cu:314-325); heading ignored.  a = row box, b = column box.
// In the reference CUDA build  a[0] - a[3]/2  is fma(a[3], -0.5, a[0]) == a[0] - a[3]*0.5 exactly, and
// Sa + Sb is contracted to fma(b[3], b[4], Sa) at every unrolled site of nms_normal_kernel (FL = 1).
template <int FL>
__device__ __forceinline__ float iou_normal(const float ax, const float ay, const float adx, const float ady,
                                            const float bx, const float by, const float bdx, const float bdy) {
    const float left = fmaxf(ax - adx * 0.5f, bx - bdx * 0.5f), right = fminf(ax + adx * 0.5f, bx + bdx * 0.5f);
    const float top = fmaxf(ay - ady * 0.5f, by - bdy * 0.5f), bottom = fminf(ay + ady * 0.5f, by + bdy * 0.5f);
    const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(w, h);
    const float sa = __fmul_rn(adx, ady);
    const float ssum = FL ? __fmaf_rn(bdx, bdy, sa) : __fadd_rn(sa, __fmul_rn(bdx, bdy));
    return __fdiv_rn(inter, fmaxf(__fsub_rn(ssum, inter), 1e-8f));
}

}  // namespace lg
