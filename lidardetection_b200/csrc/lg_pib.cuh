// lg_pib.cuh -- per-point / per-box pieces of points-in-boxes, shared by lg_points.cu and the host
// emulation of the CPU-only test tier (tests/host_emu/).
//
// Predicate (check_pt_in_box3d, /root/reference/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36),
// reproduced bit-exactly:
//     |z - cz| <= dz/2                      (closed; the reference compares in double, which for
//                                            float operands equals the float compare against dz*0.5f)
//     |lx| < dx/2 + MARGIN, |ly| < dy/2 + MARGIN   (open; compared in DOUBLE in the reference).
//       For a float v and a double D,  (double)v < D  <=>  v < RU(D)  with RU = round-up to float,
//       so the per-box thresholds are converted once (cvt.rp.f32.f64) and the per-point compare is FP32.
//     lx = fma(sx, cosa, -(sy*sina)),  ly = fma(sy, cosa, sx*sina)    with cosa = cosf(-rz), sina = sinf(-rz)
//       -- the contraction ptxas applies to lidar_to_local_coords on sm_100a (FL = 1); FL = 0 is the
//       un-contracted CPU build.
//
// Spatial cull (new): a uniform BEV grid over the frame's boxes; every cell holds a bit mask of the boxes
// whose conservative footprint touches it.  A point looks up its cell and tests only those boxes, lowest
// index first -- the reference's "first box wins" order.  The footprint bound is derived below and is
// checked against the brute-force oracle by tests/test_host_emu.py.
#pragma once
#include "lg_geom.cuh"

namespace lg {

// record: r0 = (cx, cy, cz, dz/2)   r1 = (cosa, sina, tx, ty)
template <int FL>
__device__ __forceinline__ void make_pib_record(const float* __restrict__ box, const float margin, float4& r0, float4& r1) {
    const float cx = box[0], cy = box[1], cz = box[2], dx = box[3], dy = box[4], dz = box[5], rz = box[6];
    const float cosa = trig_cos<FL>(-rz), sina = trig_sin<FL>(-rz);
    const float tx = __double2float_ru((double)dx / 2.0 + (double)margin);
    const float ty = __double2float_ru((double)dy / 2.0 + (double)margin);
    // (double)|z-cz| > (double)dz/2.0  <=>  |z-cz| > RD(dz/2); dz/2 is exact in float except for
    // subnormal underflow, where round-down keeps the equivalence.
    const float hz = __double2float_rd((double)dz / 2.0);
    r0 = make_float4(cx, cy, cz, hz);
    r1 = make_float4(cosa, sina, tx, ty);
}

template <int FL>
__device__ __forceinline__ bool pt_in_box(const float x, const float y, const float z, const float4 r0, const float4 r1) {
    if (fabsf(z - r0.z) > r0.w) return false;
    const float sx = x - r0.x, sy = y - r0.y;
    const float lx = msub<FL>(sx, r1.x, sy, r1.y);
    const float ly = madd_second<FL>(sx, r1.y, sy, r1.x);
    return (fabsf(lx) < r1.z) & (fabsf(ly) < r1.w);
}

// Conservative BEV footprint of a box: every point the predicate accepts satisfies
//     |x - cx| <= ex  and  |y - cy| <= ey.
// Proof sketch.  With c = cosa, s = sina (|c^2 + s^2 - 1| < 4e-7) and the exact lx* = sx c - sy s,
// ly* = sy c + sx s:  sx = (lx* c + ly* s)/(c^2+s^2).  The computed lx, ly differ from lx*, ly* by at most
// 2e-7 (|sx| + |sy|), and acceptance needs |lx| < tx, |ly| < ty, hence |sx| <= E + 1.2e-6 (E + F) with
// E = tx|c| + ty|s|, F = ty|c| + tx|s|; sx itself is x - cx rounded (6e-8 relative).  The slack below
// (1e-4 relative, 4e-7 |centre| for the rounding of cx +- ex) is two orders of magnitude above that.
// Returns false when the box can never contain a point (tx or ty <= 0, or a NaN anywhere);
// `bounded` is cleared when the footprint is not finite (the caller then falls back to testing every box).
__device__ __forceinline__ bool pib_footprint(const float4 r0, const float4 r1, float& ex, float& ey, bool& bounded) {
    const float c = fabsf(r1.x), s = fabsf(r1.y), tx = r1.z, ty = r1.w;
    if (!(tx > 0.f) || !(ty > 0.f) || !(c == c) || !(s == s) || !(r0.x == r0.x) || !(r0.y == r0.y)) {
        ex = ey = 0.f;
        return false;  // fabsf(l) < t is false for t <= 0 and for any NaN operand
    }
    const float E = tx * c + ty * s, F = ty * c + tx * s;
    const float slack = 1e-4f * (E + F);
    ex = E + slack + 4e-7f * fabsf(r0.x);
    ey = F + slack + 4e-7f * fabsf(r0.y);
    if (!(ex < 1e18f) || !(ey < 1e18f) || !(fabsf(r0.x) < 1e18f) || !(fabsf(r0.y) < 1e18f)) bounded = false;
    return true;
}

struct PibGrid {
    float x0, y0, invx, invy;  // cell(x) = trunc((x - x0) * invx), valid iff 0 <= (x - x0) * invx < nx
    int nx, ny;
};

// Grid over [X0, X1] x [Y0, Y1] with at most ncap cells; mean_ext = mean footprint half-extent of the boxes
// (cells much smaller than the boxes only make the marking pass longer).
__device__ __forceinline__ PibGrid pib_make_grid(float X0, float X1, float Y0, float Y1, float mean_ext, int ncap,
                                                 float min_cell_factor = 0.66f) {
    PibGrid g;
    g.x0 = X0;
    g.y0 = Y0;
    const float wx = fmaxf(X1 - X0, 1e-6f * (fabsf(X0) + fabsf(X1)) + 1e-20f);
    const float wy = fmaxf(Y1 - Y0, 1e-6f * (fabsf(Y0) + fabsf(Y1)) + 1e-20f);
    float cell = fmaxf(sqrtf(wx * wy / (float)ncap), min_cell_factor * mean_ext);
    cell = fmaxf(cell, 1e-20f);
    int nx = (int)fminf(ceilf(wx / cell), (float)ncap);
    int ny = (int)fminf(ceilf(wy / cell), (float)ncap);
    nx = max(nx, 1);
    ny = max(ny, 1);
    while ((long long)nx * ny > ncap) {  // rounding up twice can overshoot the capacity
        if (nx >= ny) nx = max(1, nx - 1);
        else ny = max(1, ny - 1);
    }
    g.nx = nx;
    g.ny = ny;
    // slightly under nx / wx so that X1 itself still maps below nx
    g.invx = (float)nx / wx * 0.99999f;
    g.invy = (float)ny / wy * 0.99999f;
    return g;
}

// the ONE mapping from a coordinate to a (fractional) cell coordinate; both the marking pass and the point
// lookup go through it, and it is monotone in v, so a point inside [lo, hi] lands in a cell inside
// [cell(lo), cell(hi)]
__device__ __forceinline__ float pib_cellf(float v, float origin, float inv) { return __fmul_rn(__fsub_rn(v, origin), inv); }

__device__ __forceinline__ int pib_cell_clamped(float v, float origin, float inv, int n) {
    const float f = pib_cellf(v, origin, inv);
    if (!(f > 0.f)) return 0;  // also NaN
    if (f >= (float)n) return n - 1;
    return (int)f;
}

// Can the footprint of box (r0, r1) reach cell (ix, iy)?  Conservative separating-axis test along the box's
// own axes (the grid axes are covered by the cell range of pib_footprint): a point q of the cell can only be
// accepted if |lx(q)| < tx, and lx(q) differs from lx(cell centre) by at most hx|c| + hy|s| (hx, hy = half
// cell size).  The cell is inflated by 1e-4 relative + 2e-6 (|x0| + extent) for the rounding of the cell
// mapping, the thresholds by the same slack as pib_footprint.
__device__ __forceinline__ bool pib_cell_touches(const float4 r0, const float4 r1, const PibGrid& g, const int ix, const int iy) {
    const float wx = 1.0f / g.invx, wy = 1.0f / g.invy;  // cell size (slightly over, see pib_make_grid)
    const float hx = 0.5f * wx * 1.0001f + 2e-6f * (fabsf(g.x0) + wx * (float)g.nx);
    const float hy = 0.5f * wy * 1.0001f + 2e-6f * (fabsf(g.y0) + wy * (float)g.ny);
    const float sx = g.x0 + ((float)ix + 0.5f) * wx - r0.x, sy = g.y0 + ((float)iy + 0.5f) * wy - r0.y;
    const float c = r1.x, s = r1.y, ac = fabsf(c), as = fabsf(s);
    const float lx = sx * c - sy * s, ly = sy * c + sx * s;
    const float E = r1.z * ac + r1.w * as, F = r1.w * ac + r1.z * as;
    const float slack = 1e-4f * (E + F) + 4e-7f * (fabsf(r0.x) + fabsf(r0.y));
    return fabsf(lx) <= r1.z + slack + hx * ac + hy * as && fabsf(ly) <= r1.w + slack + hx * as + hy * ac;
}

// The same test with everything that does not depend on the cell hoisted into a per-box constant block:
//   m0 = (cosa, sina, tX, tY)   tX = tx + slack + hx|c| + hy|s|,  tY = ty + slack + hx|s| + hy|c|
//   m1 = (x of the centre of cell column 0 minus cx, y of the centre of cell row 0 minus cy, cell width, cell height)
// so that one (box, cell) pair costs ~10 instructions.
__device__ __forceinline__ void pib_touch_consts(const float4 r0, const float4 r1, const PibGrid& g, float4& m0, float4& m1) {
    const float wx = 1.0f / g.invx, wy = 1.0f / g.invy;
    const float hx = 0.5f * wx * 1.0001f + 2e-6f * (fabsf(g.x0) + wx * (float)g.nx);
    const float hy = 0.5f * wy * 1.0001f + 2e-6f * (fabsf(g.y0) + wy * (float)g.ny);
    const float c = r1.x, s = r1.y, ac = fabsf(c), as = fabsf(s);
    const float E = r1.z * ac + r1.w * as, F = r1.w * ac + r1.z * as;
    const float slack = 1e-4f * (E + F) + 4e-7f * (fabsf(r0.x) + fabsf(r0.y));
    m0 = make_float4(c, s, r1.z + slack + hx * ac + hy * as, r1.w + slack + hx * as + hy * ac);
    m1 = make_float4(g.x0 + 0.5f * wx - r0.x, g.y0 + 0.5f * wy - r0.y, wx, wy);
}
__device__ __forceinline__ bool pib_cell_touches_fast(const float4 m0, const float4 m1, const int ix, const int iy) {
    // centre of the cell relative to the box centre; it differs from pib_cell_touches' value only by rounding,
    // which the 1e-4 relative inflation of the cell covers many times over
    const float sx = m1.x + (float)ix * m1.z, sy = m1.y + (float)iy * m1.w;
    const float lx = sx * m0.x - sy * m0.y, ly = sy * m0.x + sx * m0.y;
    return fabsf(lx) <= m0.z && fabsf(ly) <= m0.w;
}

// ---- compact candidate lists: one 32-bit word per cell ------------------------------------------
// bytes 0..3 = up to four candidate box indices in ASCENDING order, 0xFF = empty slot.  A fifth candidate turns
// byte 3 into the marker 0xFE: "more candidates exist, all with an index above byte 2" -- the point lookup then
// tests bytes 0..2 and, if none contains the point, every box above byte 2 (exact, merely slower; such cells
// are rare).  Box indices must be < 254.
constexpr uint32_t PIB_CELL_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t PIB_ID_MORE = 0xFEu, PIB_ID_NONE = 0xFFu;
constexpr int PIB_COMPACT_MAX_BOXES = 254;

__device__ __forceinline__ uint32_t pib_compact_insert(const uint32_t w, uint32_t k) {
    uint32_t a = w & 0xffu, b = (w >> 8) & 0xffu, c = (w >> 16) & 0xffu;
    const uint32_t d = w >> 24;
    uint32_t t;
    if (k < a) { t = a; a = k; k = t; }
    if (k < b) { t = b; b = k; k = t; }
    if (k < c) { t = c; c = k; k = t; }
    // a <= b <= c are the three smallest of the old (a, b, c) and k; k now holds the largest of the four
    uint32_t top;
    if (d == PIB_ID_MORE) top = PIB_ID_MORE;            // already overflowing: k (> c) joins the unnamed rest
    else if (d == PIB_ID_NONE) top = k;                  // a free slot (k may itself be the 0xFF that bubbled up)
    else top = PIB_ID_MORE;                              // five real candidates: keep the three smallest + marker
    return a | (b << 8) | (c << 16) | (top << 24);
}

}  // namespace lg
