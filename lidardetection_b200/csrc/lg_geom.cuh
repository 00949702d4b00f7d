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
//   * all per-box work (4 sincos, corners, edge vectors, margin-expanded half extents, areas, z-range,
//     volume, a conservative cull radius) is hoisted into a 112-byte record computed once per box
//     instead of once per pair (the reference re-evaluates 20 sinf/cosf per pair);
//   * the 16 corner-difference vectors q_j - p_i are formed once and every determinant is written on
//     them (negations are exact and free as operand modifiers), sharing the products the reference
//     arithmetic shares: 10 FP32 instructions per edge pair;
//   * the 16 straddle tests and 8 margin tests run uniformly into bit masks; only accepted crossings
//     are expanded, in a compact loop, and the warp is re-converged explicitly after each such loop;
//   * polygon vertices live in a per-thread shared-memory column (8 slots), not in local memory;
//   * the angular order comes from a monotone pseudo-angle packed with the vertex index into one
//     32-bit key and a 19-comparator min/max network (the reference: ~24 atan2f + bubble sort);
//   * pairs with more than 8 polygon vertices (near-coincident boxes, < 1 %), and pairs in which two
//     vertices are so close in polar angle that the reference's order hangs on the last bits of atan2f,
//     are deferred by the caller to overlap_area_slow -- a literal atan2f + stable-sort evaluation -- so
//     that they neither stall the other 31 lanes of a warp nor depart from the reference's vertex order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// sinf / cosf per arithmetic flavor (lg_trig.cuh): FL = 1 full-precision libdevice (no -use_fast_math; the CPU-only test tier
// re-points LG_SINF / LG_COSF at the oracle's restatement of libdevice when it compiles this header for the host,
// tests/host_emu/), FL = 0 glibc's double-precision algorithm restated for the device
#include "lg_trig.cuh"

namespace lg {

// ---- record layout (28 floats = 7 x float4 = 112 B) -------------------------------------------
//  rec[k], k = 0..3   (p_k.x, p_k.y, e_k.x, e_k.y): rotated corner k, order (-,-),(+,-),(+,+),(-,+), and the
//                     edge vector e_k = p_{k+1} - p_k exactly as cross() forms it (kernel.cu:34-41)
//  rec[4]             (cx, cy, cull radius, dx*dy)          -- alone feeds the cull phase
//  rec[5]             (cos(-heading), sin(-heading), dx/2 + 1e-2, dy/2 + 1e-2)   (check_in_box2d, MARGIN kernel.cu:53)
//  rec[6]             (z + dz/2, z - dz/2, dx*dy*dz, 0)     (iou3d_nms_utils.py:60-63)
constexpr int REC_F4 = 7;
constexpr int REC_FLOATS = 4 * REC_F4;
constexpr int REC_CULL = 4, REC_TRIG = 5, REC_Z = 6;
// per-lane scratch column of the fast polygon path: rows 0..7 hold the polygon's vertices, rows 8..11 the B box's corners
// (staged from the registers pair_masks loaded them into, so that the data-dependent crossing / corner loops index
// shared memory instead of re-reading the column record from global memory)
constexpr int SLAB_ROWS_SMEM_B = 8, SLAB_ROWS = 12, SLAB_BCORNER = 8;

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
// a*b - m where m is an already rounded product (shared with another determinant)
template <int FL>
__device__ __forceinline__ float msub_p(float a, float b, float m) {
    if (FL) return __fmaf_rn(a, b, -m);
    return __fsub_rn(__fmul_rn(a, b), m);
}

// ---- per-box record ----------------------------------------------------------------------------
template <int FL>
__device__ __forceinline__ void make_record_v(const float cx, const float cy, const float z, const float dx, const float dy, const float dz,
                                              const float th, float4* __restrict__ rec);

template <int FL>
__device__ __forceinline__ void make_record(const float* __restrict__ box, float4* __restrict__ rec) {
    make_record_v<FL>(box[0], box[1], box[2], box[3], box[4], box[5], box[6], rec);
}

template <int FL>
__device__ __forceinline__ void make_record_v(const float cx, const float cy, const float z, const float dx, const float dy, const float dz,
                                              const float th, float4* __restrict__ rec) {
    const float hx = __fmul_rn(dx, 0.5f), hy = __fmul_rn(dy, 0.5f);
    const float x1 = __fsub_rn(cx, hx), x2 = __fadd_rn(cx, hx), y1 = __fsub_rn(cy, hy), y2 = __fadd_rn(cy, hy);
    const float co = trig_cos<FL>(th), si = trig_sin<FL>(th);
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
    // cosf(-th), sinf(-th) of check_in_box2d (kernel.cu:55-56): both implementations are exactly even / odd -- checked over
    // every float by tests/test_host_emu.py::test_trig_of_the_negated_heading -- so the second pair of calls is a sign flip
    const float cn = co, sn = -si;
    const float tx = __fadd_rn(hx, 1e-2f), ty = __fadd_rn(hy, 1e-2f);
    const float area = __fmul_rn(dx, dy);
    // Conservative exact-zero cull.  A pair yields a polygon vertex only if an edge of
    // A crosses an edge of B, or a corner of one box lies in the other's margin-expanded rectangle; both
    // need |ca - cb| <= ra + rb with r = sqrt(tx^2 + ty^2) (tx,ty = half extents + 1 cm).  Slack on top:
    // 0.01 % + 1 mm + 1e-6 * |centre|, orders of magnitude above the few-ulp error of the rotated corners.
    const float rad = sqrtf(tx * tx + ty * ty) * 1.0001f + 1e-3f + 1e-6f * (fabsf(cx) + fabsf(cy));
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int k1 = (k + 1) & 3;
        rec[k] = make_float4(X[k], Y[k], __fsub_rn(X[k1], X[k]), __fsub_rn(Y[k1], Y[k]));
    }
    rec[REC_CULL] = make_float4(cx, cy, rad, area);
    rec[REC_TRIG] = make_float4(cn, sn, tx, ty);
    rec[REC_Z] = make_float4(__fadd_rn(z, __fmul_rn(dz, 0.5f)), __fsub_rn(z, __fmul_rn(dz, 0.5f)), __fmul_rn(area, dz), 0.f);
}

// exact-zero cull on the rec[REC_CULL] quads (cx, cy, rad, area): false => the reference returns exactly +0.0
__device__ __forceinline__ bool cull_survives(const float4 ac, const float4 bc) {
    const float dx = ac.x - bc.x, dy = ac.y - bc.y, rr = ac.z + bc.z;
    return !(dx * dx + dy * dy > rr * rr);  // NaN => keep: the polygon path decides
}

constexpr uint32_t LG_TIE_UNITS = 64;  // adjacent pseudo-angle keys closer than this are an angular near-tie

// ---- 8-input sorting network on packed 32-bit keys (19 compare-exchanges) ------------------------
__device__ __forceinline__ void cex(uint32_t& a, uint32_t& b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}
__device__ __forceinline__ void sort8(uint32_t (&k)[8]) {
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

// m | bit  iff  a0 <= b0 && a1 <= b1 && a2 <= b2 && a3 <= b3 && p1 > 0 && p2 > 0  (check_rect_cross and the two
// straddle products, kernel.cu:43-49, 75).  Written as one chain of predicate-combining compares: plain C++
// makes ptxas either branch around each edge pair ('&&') or materialise every compare in a register ('&').
__device__ __forceinline__ uint32_t or_if_crossing(uint32_t m, const uint32_t bit, float a0, float b0, float a1, float b1, float a2,
                                                   float b2, float a3, float b3, float p1, float p2) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.f32 p, %2, %3;\n\t"
        "setp.le.and.f32 p, %4, %5, p;\n\t"
        "setp.le.and.f32 p, %6, %7, p;\n\t"
        "setp.le.and.f32 p, %8, %9, p;\n\t"
        "setp.gt.and.f32 p, %10, 0f00000000, p;\n\t"
        "setp.gt.and.f32 p, %11, 0f00000000, p;\n\t"
        "@p or.b32 %0, %0, %1;\n\t}"
        : "+r"(m)
        : "r"(bit), "f"(a0), "f"(b0), "f"(a1), "f"(b1), "f"(a2), "f"(b2), "f"(a3), "f"(b3), "f"(p1), "f"(p2));
    return m;
#else
    return (a0 <= b0 && a1 <= b1 && a2 <= b2 && a3 <= b3 && p1 > 0.f && p2 > 0.f) ? (m | bit) : m;
#endif
}
// m | bit  iff  |rx| < tx && |ry| < ty   (check_in_box2d, kernel.cu:60)
__device__ __forceinline__ uint32_t or_if_inside(uint32_t m, const uint32_t bit, float rx, float tx, float ry, float ty) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t"
        "abs.f32 t, %2;\n\t"
        "setp.lt.f32 p, t, %3;\n\t"
        "abs.f32 t, %4;\n\t"
        "setp.lt.and.f32 p, t, %5, p;\n\t"
        "@p or.b32 %0, %0, %1;\n\t}"
        : "+r"(m)
        : "r"(bit), "f"(rx), "f"(tx), "f"(ry), "f"(ty));
    return m;
#else
    return (fabsf(rx) < tx && fabsf(ry) < ty) ? (m | bit) : m;
#endif
}

// ---- the pair, phase A: which of the 16 edge pairs cross, which of the 8 corners are inside ------
// A, B: records in shared (or any generic) memory.  All lanes execute the same instruction stream.
//   xmask bit 4i+j : edge i of A (p_i -> p_i+1) properly crosses edge j of B (kernel.cu:43-49, 67-75)
//   cmask bit 2k   : corner k of B lies in A's margin-expanded rectangle;  bit 2k+1: corner k of A in B's
// With D[i][j] = q_j - p_i, e_i, f_j the edge vectors (all negations below are exact):
//   s1 = cross(q0,p1,p0) = fma(Dx[i][j],  e_i.y, -(e_i.x * Dy[i][j]))
//   s2 = cross(p1,q1,p0) =  e_i.x * Dy[i][j+1]  -  Dx[i][j+1] * e_i.y         (both products rounded: they are
//                           the ones s1 of the next B edge and s5 use; ptxas keeps them unfused)
//   s3 = cross(p0,q1,q0) = fma(-Dx[i][j], f_j.y,  f_j.x * Dy[i][j])
//   s4 = cross(q1,p1,q0) = fma(f_j.x, -Dy[i+1][j], Dx[i+1][j] * f_j.y)
// A column record that lives in global memory (STAGE): the records are the only re-used global data of the sweeps, while the
// result matrix streams through the L2 once -- ask the L2 to keep them (evict_last) so that the polygon path of a late survivor does
// not find its record evicted by gigabytes of zeros.  (volatile: the lazy NMS reads records its own prologue wrote.)
__device__ __forceinline__ float4 ld_keep(const float4* p) {
#if defined(__CUDA_ARCH__) && !defined(LG_NO_EVICT_LAST)
    float4 v;
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
#else
    return *p;
#endif
}

template <int FL, bool STAGE = false>
__device__ __forceinline__ void pair_masks(const float4* __restrict__ A, const float4* __restrict__ B, uint32_t& xmask,
                                           uint32_t& cmask, float2* __restrict__ slab = nullptr, const int sstride = 0) {
    float px[4], py[4], ex[4], ey[4], qx[4], qy[4], fx[4], fy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 a = A[k], b = STAGE ? ld_keep(B + k) : B[k];
        px[k] = a.x; py[k] = a.y; ex[k] = a.z; ey[k] = a.w;
        qx[k] = b.x; qy[k] = b.y; fx[k] = b.z; fy[k] = b.w;
        if (STAGE) slab[(SLAB_BCORNER + k) * sstride] = make_float2(b.x, b.y);
    }
    // corner margin tests first (their operands die early): kernel.cu:51-61
    uint32_t cm = 0;
    {
        const float4 am = A[REC_CULL], at = A[REC_TRIG], bm = STAGE ? ld_keep(B + REC_CULL) : B[REC_CULL], bt = STAGE ? ld_keep(B + REC_TRIG) : B[REC_TRIG];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            {
                const float dx = qx[k] - am.x, dy = qy[k] - am.y;
                const float rx = msub<FL>(dx, at.x, dy, at.y);
                const float ry = madd_second<FL>(dx, at.y, dy, at.x);
                cm = or_if_inside(cm, 1u << (2 * k), rx, at.z, ry, at.w);
            }
            {
                const float dx = px[k] - bm.x, dy = py[k] - bm.y;
                const float rx = msub<FL>(dx, bt.x, dy, bt.y);
                const float ry = madd_second<FL>(dx, bt.y, dy, bt.x);
                cm = or_if_inside(cm, 1u << (2 * k + 1), rx, bt.z, ry, bt.w);
            }
        }
    }
    cmask = cm;
    // per-edge bounding intervals for check_rect_cross
    float qlox[4], qhix[4], qloy[4], qhiy[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int j1 = (j + 1) & 3;
        qlox[j] = fminf(qx[j], qx[j1]); qhix[j] = fmaxf(qx[j], qx[j1]);
        qloy[j] = fminf(qy[j], qy[j1]); qhiy[j] = fmaxf(qy[j], qy[j1]);
    }
    float Dx[4][4], Dy[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            Dx[i][j] = qx[j] - px[i];
            Dy[i][j] = qy[j] - py[i];
        }
    uint32_t xm = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int i1 = (i + 1) & 3;
        const float plox = fminf(px[i], px[i1]), phix = fmaxf(px[i], px[i1]);
        const float ploy = fminf(py[i], py[i1]), phiy = fmaxf(py[i], py[i1]);
        float m1[4], m2[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            m1[j] = __fmul_rn(ex[i], Dy[i][j]);
            m2[j] = __fmul_rn(Dx[i][j], ey[i]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int j1 = (j + 1) & 3;
            const float s1 = msub_p<FL>(Dx[i][j], ey[i], m1[j]);
            const float s2 = __fsub_rn(m1[j1], m2[j1]);
            const float s3 = FL ? __fmaf_rn(-Dx[i][j], fy[j], __fmul_rn(fx[j], Dy[i][j]))
                                : __fsub_rn(__fmul_rn(-Dx[i][j], fy[j]), __fmul_rn(fx[j], -Dy[i][j]));
            const float s4 = FL ? __fmaf_rn(fx[j], -Dy[i1][j], __fmul_rn(Dx[i1][j], fy[j]))
                                : __fsub_rn(__fmul_rn(fx[j], -Dy[i1][j]), __fmul_rn(-Dx[i1][j], fy[j]));
            xm = or_if_crossing(xm, 1u << (i * 4 + j), plox, qhix[j], qlox[j], phix, ploy, qhiy[j], qloy[j], phiy,
                                __fmul_rn(s1, s2), __fmul_rn(s3, s4));
        }
    }
    xmask = xm;
}

// crossing point of A edge i and B edge j (kernel.cu:77-91); the determinants are re-formed for this one pair
template <int FL, bool STAGE = false>
__device__ __forceinline__ float2 crossing_point(const float4* __restrict__ A, const float4* __restrict__ B, const int i,
                                                 const int j, const float2* __restrict__ slab = nullptr, const int sstride = 0) {
    const float4 pe = A[i];
    float4 q0, q1;
    if (STAGE) {  // B's corners from the lane's own scratch column; q0.z (an edge component) only in the degenerate branch below
        const float2 s0 = slab[(SLAB_BCORNER + j) * sstride], s1 = slab[(SLAB_BCORNER + ((j + 1) & 3)) * sstride];
        q0 = make_float4(s0.x, s0.y, 0.f, 0.f);
        q1 = make_float4(s1.x, s1.y, 0.f, 0.f);
    } else {
        q0 = B[j];
        q1 = B[(j + 1) & 3];
    }
    const float d0x = q0.x - pe.x, d0y = q0.y - pe.y, d1x = q1.x - pe.x, d1y = q1.y - pe.y;
    const float s1 = msub<FL>(d0x, pe.w, pe.z, d0y);
    const float t72 = __fmul_rn(pe.z, d1y), t73 = __fmul_rn(d1x, pe.w);
    const float s5 = __fsub_rn(t73, t72);  // cross(q1, p1, p0): its two products are shared with s2, unfused
    const float den = __fsub_rn(s5, s1);
    float X, Y;
    if (fabsf(den) > 1e-8f) {
        X = __fdiv_rn(msub<FL>(s5, q0.x, s1, q1.x), den);
        Y = __fdiv_rn(msub<FL>(s5, q0.y, s1, q1.y), den);
    } else {
        const float4 p1 = A[(i + 1) & 3];
        const float a0 = pe.y - p1.y, b0 = pe.z, c0 = msub<FL>(pe.x, p1.y, p1.x, pe.y);
        const float a1 = q0.y - q1.y, b1 = STAGE ? B[j].z : q0.z, c1 = msub<FL>(q0.x, q1.y, q1.x, q0.y);
        const float D = msub<FL>(a0, b1, a1, b0);
        X = __fdiv_rn(msub<FL>(b0, c1, b1, c0), D);
        Y = __fdiv_rn(msub<FL>(a1, c0, a0, c1), D);
    }
    return make_float2(X, Y);
}

// The heavy part of the fast path: 3..8 polygon vertices, from the masks of phase A.  All lanes of `hm` call it together (it
// re-converges them after each data-dependent loop).  Returns the overlap area, or -1 on an angular near-tie.
template <int FL, bool STAGE = false>
__device__ __forceinline__ float overlap_area_heavy(const float4* __restrict__ A, const float4* __restrict__ B, uint32_t xmask, uint32_t cmask,
                                                    const int cnt, float2* __restrict__ slab, const int sstride, const unsigned hm) {
    bool tie = false;
    int n = 0;
    float sx = 0.f, sy = 0.f;
    // crossing points, ascending bit order == the reference's append order (i outer, j inner)
    while (xmask) {
        const int e = __ffs(xmask) - 1;
        xmask &= xmask - 1;
        const float2 v = crossing_point<FL, STAGE>(A, B, e >> 2, e & 3, slab, sstride);
        slab[n * sstride] = v;
        sx += v.x;
        sy += v.y;
        n++;
    }
    __syncwarp(hm);
    // flagged corners (reference order: B corner k, then A corner k)
    while (cmask) {
        const int e = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        float2 c;
        if ((e & 1) || !STAGE) {
            const float4 a = (e & 1) ? A[e >> 1] : B[e >> 1];
            c = make_float2(a.x, a.y);
        } else {
            c = slab[(SLAB_BCORNER + (e >> 1)) * sstride];
        }
        slab[n * sstride] = c;
        sx += c.x;
        sy += c.y;
        n++;
    }
    __syncwarp(hm);
    const float inv = __fdividef(1.0f, (float)cnt);
    const float mx = sx * inv, my = sy * inv;  // centroid: only orders the vertices
    uint32_t key[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        key[k] = 0xFFFFFFF0u | k;
        if (k < cnt) {
            const float2 p = slab[k * sstride];
            key[k] = angle_key(p.x, p.y, mx, my, k);
        }
    }
    sort8(key);
    // near-tie of two polar angles (coincident or radially aligned vertices): the reference's order then
    // depends on atan2f's rounding -> hand the pair to the literal path.  64 key units = 1.9e-6 of the
    // pseudo-angle, ~10x the rounding noise of either angle function.
    uint32_t gap = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 7; k++)
        if (k + 1 < cnt) gap = min(gap, (key[k + 1] >> 4) - (key[k] >> 4));  // padding keys sort last
    tie = gap <= LG_TIE_UNITS;
    // fan area from the first vertex in angular order (kernel.cu:219-224); term 0 is cross(0, u) == +-0
    const float2 p0 = slab[(key[0] & 15) * sstride];
    const float2 p1 = slab[(key[1] & 15) * sstride];
    float ux = p1.x - p0.x, uy = p1.y - p0.y;
    float area = 0.f;
#pragma unroll
    for (int k = 2; k < 8; k++) {
        if (k < cnt) {
            const float2 pn = slab[(key[k] & 15) * sstride];
            const float vx = pn.x - p0.x, vy = pn.y - p0.y;
            area = __fadd_rn(area, msub<FL>(ux, vy, uy, vx));
            ux = vx;
            uy = vy;
        }
    }
    return tie ? -1.f : __fmul_rn(fabsf(area), 0.5f);
}

// ---- the pair, fast path (<= 8 polygon vertices) -------------------------------------------------
// slab: this thread's vertex column, entry k at slab[k * sstride], 8 entries.
// wmask: the lanes of this warp that call the function together (a __ballot_sync taken where the warp was
// converged); it is used to re-converge the warp after each data-dependent loop.
// Returns the overlap area exactly as the reference defines it (kernel.cu:104-225), or -1 when the polygon
// has more than 8 vertices and the pair must be handed to overlap_area_slow.
// STAGE: B lives in global memory -> its corners are staged in rows 8..11 of the lane's scratch column (the caller's slab
// then has SLAB_ROWS rows); false: B is in shared memory already, 8 rows suffice.
template <int FL, bool STAGE = false>
__device__ __forceinline__ float overlap_area(const float4* __restrict__ A, const float4* __restrict__ B,
                                              float2* __restrict__ slab, const int sstride, const unsigned wmask) {
    uint32_t xmask, cmask;
    pair_masks<FL, STAGE>(A, B, xmask, cmask, slab, sstride);
    const int cnt = __popc(xmask) + __popc(cmask);
    float res = 0.f;          // cnt <= 2: the fan sum is empty or a single zero term
    if (cnt > 8) res = -1.f;  // deferred
    const bool heavy = cnt > 2 && cnt <= 8;
    const unsigned hm = __ballot_sync(wmask, heavy);
    if (heavy) res = overlap_area_heavy<FL, STAGE>(A, B, xmask, cmask, cnt, slab, sstride, hm);
    __syncwarp(wmask);
    return res;
}

// the same as an out-of-line call: kernels whose main loop is something else (the N x M sweep: culling and storing zeros)
// keep that loop's state in registers and pay the spills only around the call
template <int FL>
__device__ __noinline__ float overlap_area_call(const float4* __restrict__ A, const float4* __restrict__ B, float2* __restrict__ slab,
                                                const int sstride, const unsigned wmask) {
    return overlap_area<FL, true>(A, B, slab, sstride, wmask);  // its callers keep B in global memory
}

// ---- the pair, 9..16 vertices (near-coincident boxes; ~0.3 % of overlapping pairs) -----------------
// Same construction as the fast path with 16 vertex slots and a 16-input network; returns -1 on an angular
// near-tie (the caller then runs the literal path).  Called for the pairs the fast path deferred.
template <int FL, typename Slab16>
__device__ __noinline__ float overlap_area16(const float4* __restrict__ A, const float4* __restrict__ B, Slab16 slab16) {
    uint32_t xmask, cmask;
    pair_masks<FL>(A, B, xmask, cmask);
    const int cnt = __popc(xmask) + __popc(cmask);  // <= 16 + 8; geometrically <= 16
    if (cnt <= 2) return 0.f;
    if (cnt > 16) return -1.f;
    int n = 0;
    float sx = 0.f, sy = 0.f;
    while (xmask) {
        const int e = __ffs(xmask) - 1;
        xmask &= xmask - 1;
        const float2 v = crossing_point<FL>(A, B, e >> 2, e & 3);
        slab16(n) = v;
        sx += v.x;
        sy += v.y;
        n++;
    }
    while (cmask) {
        const int e = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float4 c = (e & 1) ? A[e >> 1] : B[e >> 1];
        slab16(n) = make_float2(c.x, c.y);
        sx += c.x;
        sy += c.y;
        n++;
    }
    const float inv = __fdividef(1.0f, (float)cnt);
    const float mx = sx * inv, my = sy * inv;
    uint32_t key[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        key[k] = 0xFFFFFFF0u | k;
        if (k < cnt) {
            const float2 p = slab16(k);
            key[k] = angle_key(p.x, p.y, mx, my, k);
        }
    }
    sort16(key);
    uint32_t gap = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 15; k++)
        if (k + 1 < cnt) gap = min(gap, (key[k + 1] >> 4) - (key[k] >> 4));
    if (gap <= LG_TIE_UNITS) return -1.f;
    const float2 p0 = slab16((int)(key[0] & 15));
    const float2 p1 = slab16((int)(key[1] & 15));
    float ux = p1.x - p0.x, uy = p1.y - p0.y;
    float area = 0.f;
#pragma unroll
    for (int k = 2; k < 16; k++) {
        if (k < cnt) {
            const float2 pn = slab16((int)(key[k] & 15));
            const float vx = pn.x - p0.x, vy = pn.y - p0.y;
            area = __fadd_rn(area, msub<FL>(ux, vy, uy, vx));
            ux = vx;
            uy = vy;
        }
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// ---- the pair, literal path (up to 16 vertices, or angular near-ties; rare) ------------------------
// slab16(k) / ang16(k) return references to this thread's k-th vertex / angle slot, k < 16.  This is the
// reference's own procedure (kernel.cu:196-224): centroid by division, atan2f about it, stable ascending
// order (its bubble sort swaps on strict >), fan area -- written as compact rolled loops (the order is found
// by ranking), so it is small in code and free of local memory.
template <int FL, typename Slab16, typename Ang16>
__device__ __noinline__ float overlap_area_slow(const float4* __restrict__ A, const float4* __restrict__ B, Slab16 slab16,
                                                Ang16 ang16) {
    uint32_t xmask, cmask;
    pair_masks<FL>(A, B, xmask, cmask);
    const int cnt = __popc(xmask) + __popc(cmask);  // <= 16 + 8; geometrically <= 16
    if (cnt <= 2) return 0.f;
    int n = 0;
    float sx = 0.f, sy = 0.f;
    while (xmask) {
        const int e = __ffs(xmask) - 1;
        xmask &= xmask - 1;
        const float2 v = crossing_point<FL>(A, B, e >> 2, e & 3);
        if (n < 16) slab16(n) = v;
        sx = __fadd_rn(sx, v.x);
        sy = __fadd_rn(sy, v.y);
        n++;
    }
    while (cmask) {
        const int e = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float4 c = (e & 1) ? A[e >> 1] : B[e >> 1];
        if (n < 16) slab16(n) = make_float2(c.x, c.y);
        sx = __fadd_rn(sx, c.x);
        sy = __fadd_rn(sy, c.y);
        n++;
    }
    const int m = min(cnt, 16);
    const float mx = __fdiv_rn(sx, (float)cnt), my = __fdiv_rn(sy, (float)cnt);
    for (int k = 0; k < m; k++) {
        const float2 p = slab16(k);
        ang16(k) = trig_atan2<FL>(__fsub_rn(p.y, my), __fsub_rn(p.x, mx));
    }
    // order: 4 bits per rank = index of the vertex with that rank (stable: equal angles keep insertion order)
    unsigned long long order = 0ull;
    for (int k = 0; k < m; k++) {
        const float ak = ang16(k);
        int rank = 0;
        for (int l = 0; l < m; l++) {
            const float al = ang16(l);
            rank += (al < ak || (al == ak && l < k)) ? 1 : 0;
        }
        order |= (unsigned long long)k << (4 * rank);
    }
    const float2 p0 = slab16((int)(order & 15ull));
    const float2 p1 = slab16((int)((order >> 4) & 15ull));
    float ux = p1.x - p0.x, uy = p1.y - p0.y;
    float area = 0.f;
    for (int k = 2; k < m; k++) {
        const float2 pn = slab16((int)((order >> (4 * k)) & 15ull));
        const float vx = pn.x - p0.x, vy = pn.y - p0.y;
        area = __fadd_rn(area, msub<FL>(ux, vy, uy, vx));
        ux = vx;
        uy = vy;
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
