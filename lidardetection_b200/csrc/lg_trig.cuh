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

}  // namespace lg
