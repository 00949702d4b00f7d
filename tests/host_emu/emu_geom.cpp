// emu_geom.cpp -- TEST INFRASTRUCTURE: host build of the device per-pair code (see lg_host_emu.h).
#include "lg_host_emu.h"
#include "../../lidardetection_b200/csrc/lg_geom.cuh"

using namespace lg;

static long long g_slow = 0, g_literal = 0;
extern "C" long long emu_literal_count() { return g_literal; }
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
            v = overlap_area16<FL>(A, B, slab16);
            if (v < 0.f) {
                g_literal++;
                v = overlap_area_slow<FL>(A, B, slab16, ang16);
            }
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

// ---- points in boxes: the grid algorithm of lg_points.cu, serial (same lg_pib.cuh functions) ----------------
#include "../../lidardetection_b200/csrc/lg_pib.cuh"

#include <vector>

// boxes (T,7), pts (M,3) -> out (M); ncells = capacity of the cell array (kernel: 4096).
// returns the number of (point, box) predicate evaluations performed (for the pruning statistics)
extern "C" long long emu_points_in_boxes(const float* boxes, int T, const float* pts, long long M, int32_t* out, int ncells,
                                         int* used_grid) {
    std::vector<float4> rec(2 * (size_t)T);
    float lo_x = INFINITY, hi_x = -INFINITY, lo_y = INFINITY, hi_y = -INFINITY, sum_ext = 0.f, nv = 0.f;
    bool bounded = true;
    for (int k = 0; k < T; k++) {
        make_pib_record<1>(boxes + k * 7, 1e-5f, rec[2 * k], rec[2 * k + 1]);
        float ex, ey;
        if (pib_footprint(rec[2 * k], rec[2 * k + 1], ex, ey, bounded)) {
            lo_x = fminf(lo_x, rec[2 * k].x - ex); hi_x = fmaxf(hi_x, rec[2 * k].x + ex);
            lo_y = fminf(lo_y, rec[2 * k].y - ey); hi_y = fmaxf(hi_y, rec[2 * k].y + ey);
            sum_ext += 0.5f * (ex + ey);
            nv += 1.f;
        }
    }
    const bool use_grid = bounded && T <= PIB_COMPACT_MAX_BOXES;
    *used_grid = use_grid ? 1 : 0;
    long long tests = 0;
    std::vector<uint32_t> cells((size_t)ncells, PIB_CELL_EMPTY);
    PibGrid g{};
    if (use_grid && nv > 0.f) {
        g = pib_make_grid(lo_x, hi_x, lo_y, hi_y, sum_ext / nv, ncells, 0.6f);
        for (int k = T - 1; k >= 0; k--) {  // any insertion order must give the same lists: go backwards on purpose
            float ex, ey;
            bool dummy = true;
            const float4 r0 = rec[2 * k], r1 = rec[2 * k + 1];
            if (!pib_footprint(r0, r1, ex, ey, dummy)) continue;
            const int ix0 = pib_cell_clamped(r0.x - ex, g.x0, g.invx, g.nx), ix1 = pib_cell_clamped(r0.x + ex, g.x0, g.invx, g.nx);
            const int iy0 = pib_cell_clamped(r0.y - ey, g.y0, g.invy, g.ny), iy1 = pib_cell_clamped(r0.y + ey, g.y0, g.invy, g.ny);
            float4 m0, m1;
            pib_touch_consts(r0, r1, g, m0, m1);
            for (int iy = iy0; iy <= iy1; iy++)
                for (int ix = ix0; ix <= ix1; ix++)
                    if (pib_cell_touches_fast(m0, m1, ix, iy)) cells[(size_t)iy * g.nx + ix] = pib_compact_insert(cells[(size_t)iy * g.nx + ix], (uint32_t)k);
        }
    }
    for (long long p = 0; p < M; p++) {
        const float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
        int r = -1;
        if (use_grid) {
            const int ix = __float2int_rd(pib_cellf(x, g.x0, g.invx)), iy = __float2int_rd(pib_cellf(y, g.y0, g.invy));
            if (nv > 0.f && (unsigned)ix < (unsigned)g.nx && (unsigned)iy < (unsigned)g.ny) {
                const uint32_t ids = cells[(size_t)iy * g.nx + ix];
                for (int sl = 0; sl < 4 && r < 0; sl++) {
                    const uint32_t id = (ids >> (8 * sl)) & 0xffu;
                    if (id >= PIB_ID_MORE) break;
                    tests++;
                    if (pt_in_box<1>(x, y, z, rec[2 * id], rec[2 * id + 1])) r = (int)id;
                }
                if (r < 0 && (ids >> 24) == PIB_ID_MORE)
                    for (int k = (int)((ids >> 16) & 0xffu) + 1; k < T; k++) {
                        tests++;
                        if (pt_in_box<1>(x, y, z, rec[2 * k], rec[2 * k + 1])) { r = k; break; }
                    }
            }
        } else {
            for (int k = 0; k < T; k++) {
                tests++;
                if (pt_in_box<1>(x, y, z, rec[2 * k], rec[2 * k + 1])) { r = k; break; }
            }
        }
        out[p] = r;
    }
    return tests;
}

// ---- the device restatement of glibc's sinf / cosf (lg_trig.cuh), for the sweep against the host's libm ----
extern "C" void emu_glibc_sincosf(const float* x, long long n, float* out, int want_cos) {
    for (long long i = 0; i < n; i++) out[i] = glibc_sincosf(x[i], want_cos);
}
// sweep over float bit patterns first, first + stride, ... (count of them); returns the number of results whose bits differ
// from the host libm's sinf / cosf (NaN results compare equal to each other)
extern "C" long long emu_glibc_sweep(unsigned first, unsigned stride, long long count, unsigned* first_bad) {
    long long bad = 0;
    unsigned u = first;
    for (long long i = 0; i < count; i++, u += stride) {
        const float f = __uint_as_float(u);
        const float s0 = sinf(f), s1 = glibc_sincosf(f, 0), c0 = cosf(f), c1 = glibc_sincosf(f, 1);
        const bool bs = __float_as_uint(s0) != __float_as_uint(s1) && !(s0 != s0 && s1 != s1);
        const bool bc = __float_as_uint(c0) != __float_as_uint(c1) && !(c0 != c0 && c1 != c1);
        if ((bs || bc) && bad++ == 0 && first_bad) *first_bad = u;
    }
    return bad;
}

// all-pairs 0/1 mask of points_in_boxes_cpu (pib_mask_kernel), flavor 0 = the reference CPU build
extern "C" void emu_points_mask(const float* boxes, int T, const float* pts, long long M, int32_t* out, int flavor) {
    for (int k = 0; k < T; k++) {
        float4 r0, r1;
        if (flavor) make_pib_record<1>(boxes + k * 7, 1e-5f, r0, r1);
        else make_pib_record<0>(boxes + k * 7, 1e-5f, r0, r1);
        for (long long p = 0; p < M; p++) {
            const float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
            out[(long long)k * M + p] = (flavor ? pt_in_box<1>(x, y, z, r0, r1) : pt_in_box<0>(x, y, z, r0, r1)) ? 1 : 0;
        }
    }
}

// sinf(-x) == -sinf(x) and cosf(-x) == cosf(x), bit for bit, in both flavors' implementations (make_record computes the
// trigonometry of -heading from that of +heading): count of float bit patterns first, first + stride, ... that violate it
extern "C" long long emu_trig_symmetry_sweep(unsigned first, unsigned stride, long long count, int flavor, unsigned* first_bad) {
    long long bad = 0;
    unsigned u = first;
    for (long long i = 0; i < count; i++, u += stride) {
        const float f = __uint_as_float(u & 0x7fffffffu);
        if (!(f == f)) continue;
        const float sp = flavor ? lgo_sinf(f, 1) : glibc_sincosf(f, 0), sn = flavor ? lgo_sinf(-f, 1) : glibc_sincosf(-f, 0);
        const float cp = flavor ? lgo_cosf(f, 1) : glibc_sincosf(f, 1), cn = flavor ? lgo_cosf(-f, 1) : glibc_sincosf(-f, 1);
        const bool ok = (__float_as_uint(sn) == (__float_as_uint(sp) ^ 0x80000000u) || (sp != sp && sn != sn)) &&
                        (__float_as_uint(cn) == __float_as_uint(cp) || (cp != cp && cn != cn));
        if (!ok && bad++ == 0 && first_bad) *first_bad = u;
    }
    return bad;
}

// the device restatement of glibc's atan2f against the host libm: count of differing results over n operand pairs
extern "C" long long emu_glibc_atan2f_check(const float* y, const float* x, long long n) {
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        const float a = atan2f(y[i], x[i]), b = glibc_atan2f(y[i], x[i]);
        if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) bad++;
    }
    return bad;
}
