// lg_trig.cuh -- sinf / cosf of the reference's two builds, selectable per arithmetic flavor.
//
//   FL = 1  reference CUDA build: libdevice sinf / cosf (full precision, no fast-math).
//   FL = 0  reference CPU build (iou3d_cpu.cpp:94-98, 169; roiaware_pool3d.cpp:121-140): glibc's sinf / cosf.
//           The host's libm cannot be called from a kernel, and libdevice's results differ from it in the last
//           bit for ~4 % of the headings -- enough to move an IoU of small boxes at range by 3e-5.  glibc >= 2.28
//           computes both functions in DOUBLE precision (range reduction by pi/2 with one fused multiply-subtract,
//           degree-7 / degree-8 minimax polynomials, one final rounding to float); that algorithm is deterministic
//           and is restated here operation by operation, so the *_cpu entry points of the drop-in return the
//           reference CPU build's bits without a host fallback.
//           Published algorithm: glibc 2.39 sysdeps/ieee754/flt-32/{s_sinf.c, s_cosf.c, s_sincosf.h, s_sincosf_data.c}
//           (ARM Optimized Routines sincosf), x86-64 ifunc variant `__sinf_fma` / `__cosf_fma` (built with -mfma:
//           every a*b + c of the source is one fused operation).  tests/test_host_emu.py compiles this header for the
//           host and checks it against the libm of the container over a 2^28-point sweep of all float bit patterns
//           (tools/check_glibc_trig.py is the exhaustive 2^32 version: 0 mismatches for sinf and cosf on glibc 2.39).
//           A host without FMA would take the `_sse2` variant, whose results differ on 34 of the 2^32 inputs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LG_SINF
#define LG_SINF(x) sinf(x)
#define LG_COSF(x) cosf(x)
#endif
#ifndef LG_ATAN2F
#define LG_ATAN2F(y, x) atan2f((y), (x))
#endif

namespace lg {

// 4/pi as a bit string, 24 overlapping 32-bit windows (s_sincosf_data.c: __inv_pio4)
__device__ __forceinline__ uint32_t inv_pio4_word(int i) {
    constexpr uint32_t W[24] = {0xa2u,       0xa2f9u,     0xa2f983u,   0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
                                0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
                                0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};
    return W[i];
}

// want_cos = 0: sinf(y), 1: cosf(y).  Bit-identical to glibc's FMA build for every float (NaN payloads aside).
static __device__ __noinline__ float glibc_sincosf(const float y, const int want_cos) {
    const uint32_t yi = __float_as_uint(y);
    const uint32_t top = (yi >> 20) & 0x7ffu;  // abstop12
    double x = (double)y;
    int n, flip = 0;  // flip: the second coefficient table = the first with the cosine polynomial negated
    if (top < 0x3f4u) {                                 // |y| < pi/4
        if (top < 0x398u) return want_cos ? 1.0f : y;  // |y| < 2^-12
        n = want_cos;
    } else {
        int q;
        if (top < 0x42fu) {  // |y| < 120: reduce_fast -- n = round(x * 2/pi) in 8.24 fixed point, x -= n * pi/2 (fused)
            const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
            q = (__double2int_rz(r) + 0x800000) >> 24;
            x = __fma_rn(-(double)q, 0x1.921FB54442D18p0, x);
            n = q;
        } else if (top < 0x7f8u) {  // reduce_large: exact 32 x 96 -> 128-bit fixed-point product with 4/pi
            const int w = (int)((yi >> 26) & 15u), shift = (int)((yi >> 23) & 7u);
            const uint32_t m = ((yi & 0xffffffu) | 0x800000u) << shift;
            unsigned long long res0 = (unsigned long long)(uint32_t)(m * inv_pio4_word(w));
            const unsigned long long res1 = (unsigned long long)m * inv_pio4_word(w + 4);
            const unsigned long long res2 = (unsigned long long)m * inv_pio4_word(w + 8);
            res0 = (res2 >> 32) | (res0 << 32);
            res0 += res1;
            const unsigned long long nn = (res0 + (1ull << 61)) >> 62;
            res0 -= nn << 62;
            x = __dmul_rn(__ll2double_rn((long long)res0), 0x1.921FB54442D18p-62);
            n = (int)nn;
            q = n + (int)(yi >> 31);  // the sign of y joins the quadrant
        } else {
            return __fsub_rn(y, y);  // inf / NaN -> NaN
        }
        if ((q + 1) & 2) x = -x;     // sign[q & 3] = {1, -1, -1, 1}
        flip = (q & 2) != 0;
        n ^= want_cos;
    }
    const double x2 = __dmul_rn(x, x);
    if ((n & 1) == 0) {  // sine polynomial (identical in both tables)
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = __dmul_rn(x3, x2);
        const double s = __fma_rn(x3, -0x1.555545995a603p-3, x);
        return __double2float_rn(__fma_rn(x7, s1, s));
    }
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, 0x1.99343027bf8c3p-16, -0x1.6c087e89a359dp-10);
    const double c1 = __fma_rn(x2, -0x1.ffffffd0c621cp-2, 1.0);
    const double x6 = __dmul_rn(x4, x2);
    const double c = __fma_rn(x4, 0x1.55553e1068f19p-5, c1);
    const float r = __double2float_rn(__fma_rn(x6, c2, c));
    return flip ? -r : r;  // round-to-nearest is symmetric: negating every coefficient negates the result exactly
}

// glibc's atan2f / atanf (2.39: sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c -- the fdlibm single-precision code, every
// operation individually rounded; no FMA variant exists for it), restated for the device: the literal vertex ordering of the
// reference CPU build (iou3d_cpu.cpp:196-211) hangs on its last bit when two polygon vertices are radially aligned.
// Checked against the container's libm on 4e8 operand pairs (random bit patterns and coordinate-like values): 0 differences;
// tests/test_host_emu.py repeats a shorter sweep.  Infinite operands are not special-cased (the callers pass finite differences).
static __device__ __noinline__ float glibc_atanf(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const int hx = (int)__float_as_uint(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000) return __fadd_rn(x, x);
        const float r = __fadd_rn(atanhi[3], atanlo[3]);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {      // |x| < 1.1875
            if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
                id = 0;
                x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x));
            } else {
                id = 1;
                x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f));
            }
        } else if (ix < 0x401c0000) {  // |x| < 2.4375
            id = 2;
            x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x)));
        } else {
            id = 3;
            x = __fdiv_rn(-1.0f, x);
        }
    }
    const float z = __fmul_rn(x, x), w = __fmul_rn(z, z);
    auto h = [](float a, float b, float c) { return __fadd_rn(a, __fmul_rn(b, c)); };  // a + b*c, two roundings
    const float s1 = __fmul_rn(z, h(3.3333334327e-01f, w, h(1.4285714924e-01f, w, h(9.0908870101e-02f, w, h(6.6610731184e-02f, w, h(4.9768779427e-02f, w, 1.6285819933e-02f))))));
    const float s2 = __fmul_rn(w, h(-2.0000000298e-01f, w, h(-1.1111110449e-01f, w, h(-7.6918758452e-02f, w, h(-5.8335702866e-02f, w, -3.6531571299e-02f)))));
    const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
    if (id < 0) return __fsub_rn(x, xs);
    const float r = __fsub_rn(atanhi[id], __fsub_rn(__fsub_rn(xs, atanlo[id]), x));
    return hx < 0 ? -r : r;
}

static __device__ __noinline__ float glibc_atan2f(const float y, const float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int hx = (int)__float_as_uint(x), hy = (int)__float_as_uint(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
    if (hx == 0x3f800000) return glibc_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2 * sign(x) + sign(y)
    if (iy == 0) return m < 2 ? y : (m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny));
    if (ix == 0) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = glibc_atanf(fabsf(__fdiv_rn(y, x)));
    if (m == 0) return z;
    if (m == 1) return -z;
    if (m == 2) return __fsub_rn(pi, __fsub_rn(z, pi_lo));
    return __fsub_rn(__fsub_rn(z, pi_lo), pi);
}

template <int FL>
__device__ __forceinline__ float trig_sin(const float x) {
    if (FL) return LG_SINF(x);
    return glibc_sincosf(x, 0);
}
template <int FL>
__device__ __forceinline__ float trig_cos(const float x) {
    if (FL) return LG_COSF(x);
    return glibc_sincosf(x, 1);
}

template <int FL>
__device__ __forceinline__ float trig_atan2(const float y, const float x) {
    if (FL) return LG_ATAN2F(y, x);
    return glibc_atan2f(y, x);
}

}  // namespace lg
