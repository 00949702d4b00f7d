// lg_host_emu.h -- TEST INFRASTRUCTURE: lets g++ compile the per-pair device headers of
// lidardetection_b200/csrc (lg_geom.cuh, lg_pib.cuh) for the HOST, so that their arithmetic can be
// compared bit-for-bit with the oracle in the CPU-only test tier (no GPU in the dev container).
// Every CUDA intrinsic used by those headers is mapped to the IEEE operation it denotes; build with
// -ffp-contract=off so that nothing is re-contracted.  sinf/cosf are redirected to the oracle's
// restatement of CUDA libdevice (oracle/lg_oracle.c: lgo_sinf / lgo_cosf, flavor 1).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

extern "C" float lgo_sinf(float, int);
extern "C" float lgo_cosf(float, int);
#define LG_SINF(x) lgo_sinf((x), 1)
#define LG_COSF(x) lgo_cosf((x), 1)

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

using std::max;
using std::min;

static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }  // approximate on the device: ordering keys only
static inline float __frcp_rn(float a) { return 1.0f / a; }
// double-precision intrinsics of lg_trig.cuh (the restatement of glibc's sinf / cosf)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __double2int_rz(double a) { return (int)a; }
static inline double __ll2double_rn(long long a) { return (double)a; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline unsigned __float2uint_rn(float x) { return (unsigned)llrintf(x); }
static inline int __float2int_rd(float x) { return (int)floorf(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline float __double2float_ru(double d) {
    float f = (float)d;
    if ((double)f < d) f = nextafterf(f, INFINITY);
    return f;
}
static inline float __double2float_rd(double d) {
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}
static inline unsigned __ballot_sync(unsigned m, bool p) { return (m & 1u) && p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __activemask() { return 1u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
