// lg_kitti.cu -- SURVEY 8f-2: the KITTI evaluation's rotated overlaps.
//
// Replaces, behind the C ABI,
//   rotate_iou_gpu_eval / rotate_iou_kernel_eval   pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:260-330 (numba.cuda)
//   bev_box_overlap, d3_box_overlap(+_kernel)      pcdet/datasets/kitti/kitti_object_eval_python/eval.py:111-155
//   the per-part loop of calculate_iou_partly      eval.py:340-414 (one H2D + launch + D2H + numba-CPU pass per part)
//
// Conventions of that code (NOT those of iou3d_nms): 5-parameter boxes (cx, cy, x_d, y_d, angle), angle clockwise, no
// margin, closed containment, iou[n, k] = f(query_boxes[k], boxes[n]), `criterion` -1 / 0 / 1 / other.
//
// Arithmetic contract: the operation order, operand types (numba types python float literals as float64) and FMA
// contraction of the reference kernel as numba 0.65 -> NVVM -> ptxas 12.9 build it for sm_100a; see the block comment in
// oracle/lg_oracle.c ("Next row 8f-2") for the list, read off that build's PTX / SASS.  LG_FLAG_STRICT_FP32 evaluates the
// same statements without contraction.  Every float operation below is an explicit _rn intrinsic: nvcc never re-contracts
// those.
//
// Structure: a prep kernel turns every box into a 48-byte record (rotated corners, centre, circum-radius, area) -- the
// reference recomputes sin/cos and the corners for every PAIR.  The pair kernel gives every 256-thread CTA one 32-row x
// 256-column tile of one part's matrix (column constants in registers, row constants broadcast from shared memory,
// stores coalesced along the columns): phase 1 puts every pair to an exact-zero test (circum-circles apart by a slack
// that dwarfs any rounding of the reference's predicates => the reference collects < 3 polygon points => its area is
// exactly 0.0) and stores those results; the survivors (~0.5 % of a KITTI part, ~40 of a tile's 8192 pairs) are queued
// in shared memory and phase 2 runs the polygon path on the queue.  Several evaluation "parts" (independent N_p x K_p
// problems, eval.py:356-395) go into one launch: a one-CTA kernel scans the parts' tile counts, the pair kernel's CTAs
// find their part by binary search over that prefix.
#include "lg_common.cuh"

namespace lg {
namespace kitti {

constexpr int THREADS = 256;
constexpr int MAXPTS = 24;  // 8 corners + 16 crossings; the reference's buffer holds 8 (undefined beyond, see oracle)

struct Rec {        // 48 bytes
    float c[8];     // rotated corners x0 y0 .. x3 y3 (rbbox_to_corners)
    float cx, cy;   // centre
    float r;        // half diagonal
    float area;     // x_d * y_d
};
struct RecZ {       // d3_box_overlap: float64 camera-frame extras
    double y, ybot, vol;
};

template <int FL>
__device__ __forceinline__ float msub(float a, float b, float c, float d) {  // a*b - c*d
    if (FL) return __fmaf_rn(a, b, -__fmul_rn(c, d));
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
template <int FL>
__device__ __forceinline__ float madd(float a, float b, float c, float d) {  // a*b + c*d, first product fused
    if (FL) return __fmaf_rn(a, b, __fmul_rn(c, d));
    return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}

// rbbox_to_corners (rotate_iou.py:201-227) + the per-box constants
template <int FL>
__device__ __forceinline__ void make_rec(float bx, float by, float xd, float yd, float ang, Rec& R) {
    const float a_cos = cosf(ang), a_sin = sinf(ang);
    const float cxs[4] = {__fmul_rn(xd, -0.5f), __fmul_rn(xd, -0.5f), __fmul_rn(xd, 0.5f), __fmul_rn(xd, 0.5f)};
    const float cys[4] = {__fmul_rn(yd, -0.5f), __fmul_rn(yd, 0.5f), __fmul_rn(yd, 0.5f), __fmul_rn(yd, -0.5f)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (FL) {
            R.c[2 * i] = __fadd_rn(bx, __fmaf_rn(a_cos, cxs[i], __fmul_rn(a_sin, cys[i])));
            R.c[2 * i + 1] = __fadd_rn(by, __fmaf_rn(a_cos, cys[i], -__fmul_rn(a_sin, cxs[i])));
        } else {
            R.c[2 * i] = __fadd_rn(__fadd_rn(__fmul_rn(a_cos, cxs[i]), __fmul_rn(a_sin, cys[i])), bx);
            R.c[2 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(-a_sin, cxs[i]), __fmul_rn(a_cos, cys[i])), by);
        }
    }
    R.cx = bx;
    R.cy = by;
    R.r = 0.5f * sqrtf(xd * xd + yd * yd);
    // A quadrilateral with a zero-length side (dims of 0, or below the ulp of the centre) makes point_in_quadrilateral's
    // tests along that side read 0 >= 0 >= 0: it "contains" an infinite strip, so distance proves nothing.  Such a box is
    // never culled (infinite radius => the exact-zero test is false for every pair it takes part in).
    const float e0 = __fsub_rn(R.c[2], R.c[0]), e1 = __fsub_rn(R.c[3], R.c[1]), f0 = __fsub_rn(R.c[6], R.c[0]), f1 = __fsub_rn(R.c[7], R.c[1]);
    if (!(e0 * e0 + e1 * e1 >= 1e-30f && f0 * f0 + f1 * f1 >= 1e-30f)) R.r = __int_as_float(0x7f800000);
    R.area = __fmul_rn(xd, yd);
}

// in_f32 != nullptr: rows of 5 floats; else in_f64: rows of 7 doubles (x, y, z, l, h, w, ry) -> columns [0, 2, 3, 5, 6]
template <int FL>
__global__ void __launch_bounds__(256) kitti_prep_kernel(const float* __restrict__ in_f32, const double* __restrict__ in_f64, int64_t n,
                                                         Rec* __restrict__ rec, RecZ* __restrict__ recz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float bx, by, xd, yd, ang;
    if (in_f32) {
        const float* b = in_f32 + 5 * i;
        bx = b[0], by = b[1], xd = b[2], yd = b[3], ang = b[4];
    } else {
        const double* b = in_f64 + 7 * i;
        bx = __double2float_rn(b[0]), by = __double2float_rn(b[2]), xd = __double2float_rn(b[3]), yd = __double2float_rn(b[5]);
        ang = __double2float_rn(b[6]);
        if (recz) {
            RecZ z;
            z.y = b[1];
            z.ybot = __dsub_rn(b[1], b[4]);
            z.vol = __dmul_rn(__dmul_rn(b[3], b[4]), b[5]);
            recz[i] = z;
        }
    }
    Rec R;
    make_rec<FL>(bx, by, xd, yd, ang, R);
    float4* o = reinterpret_cast<float4*>(rec + i);
    o[0] = make_float4(R.c[0], R.c[1], R.c[2], R.c[3]);
    o[1] = make_float4(R.c[4], R.c[5], R.c[6], R.c[7]);
    o[2] = make_float4(R.cx, R.cy, R.r, R.area);
}

// point_in_quadrilateral (rotate_iou.py:155-173)
template <int FL>
__device__ __forceinline__ bool in_quad(float px, float py, const float* c) {
    const float ab0 = __fsub_rn(c[2], c[0]), ab1 = __fsub_rn(c[3], c[1]), ad0 = __fsub_rn(c[6], c[0]), ad1 = __fsub_rn(c[7], c[1]);
    const float ap0 = __fsub_rn(px, c[0]), ap1 = __fsub_rn(py, c[1]);
    float abab, abap, adad, adap;
    if (FL) {
        abab = __fmaf_rn(ab0, ab0, __fmul_rn(ab1, ab1));
        abap = __fmaf_rn(ab1, ap1, __fmul_rn(ab0, ap0));
        adad = __fmaf_rn(ad0, ad0, __fmul_rn(ad1, ad1));
        adap = __fmaf_rn(ad1, ap1, __fmul_rn(ad0, ap0));
    } else {
        abab = __fadd_rn(__fmul_rn(ab0, ab0), __fmul_rn(ab1, ab1));
        abap = __fadd_rn(__fmul_rn(ab0, ap0), __fmul_rn(ab1, ap1));
        adad = __fadd_rn(__fmul_rn(ad0, ad0), __fmul_rn(ad1, ad1));
        adap = __fadd_rn(__fmul_rn(ad0, ap0), __fmul_rn(ad1, ap1));
    }
    return abab >= abap && abap >= 0.0f && adad >= adap && adap >= 0.0f;
}

// inter() (rotate_iou.py:230-243): quadrilateral_intersection + sort_vertex_in_convex_polygon + area; c1 = first argument
template <int FL>
__device__ __noinline__ double inter_area(const float* __restrict__ c1, const float* __restrict__ c2) {
    float pts[2 * MAXPTS], vs[MAXPTS];
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (in_quad<FL>(c1[2 * i], c1[2 * i + 1], c2)) {
            pts[2 * n] = c1[2 * i], pts[2 * n + 1] = c1[2 * i + 1];
            ++n;
        }
        if (in_quad<FL>(c2[2 * i], c2[2 * i + 1], c1)) {
            pts[2 * n] = c2[2 * i], pts[2 * n + 1] = c2[2 * i + 1];
            ++n;
        }
    }
    // line_segment_intersection (rotate_iou.py:77-116), A->B an edge of c1, C->D an edge of c2
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float A0 = c1[2 * i], A1 = c1[2 * i + 1], B0 = c1[2 * ((i + 1) & 3)], B1 = c1[2 * ((i + 1) & 3) + 1];
        const float BA0 = __fsub_rn(B0, A0), BA1 = __fsub_rn(B1, A1);
        const float ABBA = msub<FL>(A0, B1, B0, A1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float C0 = c2[2 * j], C1 = c2[2 * j + 1], D0 = c2[2 * ((j + 1) & 3)], D1 = c2[2 * ((j + 1) & 3) + 1];
            const float DA0 = __fsub_rn(D0, A0), CA0 = __fsub_rn(C0, A0), DA1 = __fsub_rn(D1, A1), CA1 = __fsub_rn(C1, A1);
            const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
            const bool bcd = __fmul_rn(__fsub_rn(D1, B1), __fsub_rn(C0, B0)) > __fmul_rn(__fsub_rn(C1, B1), __fsub_rn(D0, B0));
            if (acd != bcd) {
                const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
                const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
                if (abc != abd) {
                    const float DC0 = __fsub_rn(D0, C0), DC1 = __fsub_rn(D1, C1);
                    const float CDDC = msub<FL>(C0, D1, D0, C1);
                    const float DH = msub<FL>(BA1, DC0, BA0, DC1);
                    const float Dx = msub<FL>(ABBA, DC0, BA0, CDDC);
                    const float Dy = msub<FL>(ABBA, DC1, BA1, CDDC);
                    pts[2 * n] = __fdiv_rn(Dx, DH), pts[2 * n + 1] = __fdiv_rn(Dy, DH);
                    ++n;
                }
            }
        }
    }
    if (n > 0) {  // sort_vertex_in_convex_polygon (rotate_iou.py:36-74)
        float s0 = 0.0f, s1 = 0.0f;
        for (int i = 0; i < n; ++i) s0 = __fadd_rn(s0, pts[2 * i]), s1 = __fadd_rn(s1, pts[2 * i + 1]);
        const double dn = (double)n;
        const float m0 = __double2float_rn(__ddiv_rn((double)s0, dn)), m1 = __double2float_rn(__ddiv_rn((double)s1, dn));
        for (int i = 0; i < n; ++i) {
            float v0 = __fsub_rn(pts[2 * i], m0), v1 = __fsub_rn(pts[2 * i + 1], m1);
            const float d = __fsqrt_rn(FL ? __fmaf_rn(v0, v0, __fmul_rn(v1, v1)) : __fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)));
            v0 = __fdiv_rn(v0, d);
            v1 = __fdiv_rn(v1, d);
            if (v1 < 0.0f) v0 = __fsub_rn(-2.0f, v0);
            vs[i] = v0;
        }
        for (int i = 1; i < n; ++i) {
            if (vs[i - 1] > vs[i]) {
                const float temp = vs[i], tx = pts[2 * i], ty = pts[2 * i + 1];
                int j = i;
                while (j > 0 && vs[j - 1] > temp) {
                    vs[j] = vs[j - 1];
                    pts[2 * j] = pts[2 * j - 2];
                    pts[2 * j + 1] = pts[2 * j - 1];
                    --j;
                }
                vs[j] = temp;
                pts[2 * j] = tx;
                pts[2 * j + 1] = ty;
            }
        }
    }
    double area = 0.0;  // area() / trangle_area() (rotate_iou.py:17-33): float cross product, halved / abs / summed in double
    for (int i = 0; i < n - 2; ++i) {
        const float a0 = pts[0], a1 = pts[1], b0 = pts[2 * i + 2], b1 = pts[2 * i + 3], c0 = pts[2 * i + 4], cc1 = pts[2 * i + 5];
        float t;
        if (FL)
            t = __fmaf_rn(__fsub_rn(a0, c0), __fsub_rn(b1, cc1), -__fmul_rn(__fsub_rn(a1, cc1), __fsub_rn(b0, c0)));
        else
            t = __fsub_rn(__fmul_rn(__fsub_rn(a0, c0), __fsub_rn(b1, cc1)), __fmul_rn(__fsub_rn(a1, cc1), __fsub_rn(b0, c0)));
        area = __dadd_rn(area, fabs(__dmul_rn((double)t, 0.5)));
    }
    return area;
}

// devRotateIoUEval (rotate_iou.py:246-258); area1 belongs to the FIRST argument (the query box)
__device__ __forceinline__ float finish_bev(double ai, float area1, float area2, int criterion) {
    double r;
    if (criterion == -1)
        r = __ddiv_rn(ai, __dsub_rn((double)__fadd_rn(area1, area2), ai));
    else if (criterion == 0)
        r = __ddiv_rn(ai, (double)area1);
    else if (criterion == 1)
        r = __ddiv_rn(ai, (double)area2);
    else
        r = ai;
    return __double2float_rn(r);
}

// finish_bev(0.0, ...) without the double division: 0.0 / x is +0 for x > 0, -0 for x < 0 and NaN for x == 0 or NaN
// (float -> double is exact, so the sign test on the float operand decides the same way)
__device__ __forceinline__ float zero_bev(float area1, float area2, int criterion) {
    float x;
    if (criterion == -1)
        x = __fadd_rn(area1, area2);
    else if (criterion == 0)
        x = area1;
    else if (criterion == 1)
        x = area2;
    else
        return 0.0f;
    return x > 0.0f ? 0.0f : (x < 0.0f ? -0.0f : __int_as_float(0x7fffffff));
}

// d3_box_overlap_kernel (eval.py:116-147) applied to rinc = float32(BEV intersection area); b = box, q = query box
__device__ __forceinline__ float finish_d3(float rinc, const RecZ& b, const RecZ& q, int criterion) {
    if (!(rinc > 0.0f)) return rinc;
    const double iw = __dsub_rn(fmin(b.y, q.y), fmax(b.ybot, q.ybot));
    if (!(iw > 0.0)) return 0.0f;
    const double inc = __dmul_rn(iw, (double)rinc);
    double ua;
    if (criterion == -1)
        ua = __dsub_rn(__dadd_rn(b.vol, q.vol), inc);
    else if (criterion == 0)
        ua = b.vol;
    else if (criterion == 1)
        ua = q.vol;
    else
        ua = inc;
    return __double2float_rn(__ddiv_rn(inc, ua));
}

struct Parts {
    const int64_t* box_off;  // [P + 1] or nullptr (single part)
    const int64_t* q_off;
    const int64_t* out_off;
    const int64_t* tile_prefix;  // [P + 1], written by kitti_tiles_kernel (parts only)
    int num_parts;
    int64_t n, k;  // single part
};

constexpr int TR = 32;        // rows (boxes) per tile
constexpr int TC = THREADS;   // columns (query boxes) per tile: one per thread

__host__ __device__ inline int64_t tiles_of(int64_t n, int64_t k) { return ((n + TR - 1) / TR) * ((k + TC - 1) / TC); }

// exclusive prefix of the parts' tile counts (one CTA; parts are few: 51 for a KITTI val evaluation)
__global__ void __launch_bounds__(256) kitti_tiles_kernel(const int64_t* __restrict__ box_off, const int64_t* __restrict__ q_off,
                                                          int num_parts, int64_t* __restrict__ tile_prefix) {
    __shared__ int64_t s_scan[256];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int p0 = 0; p0 < num_parts; p0 += 256) {
        const int p = p0 + threadIdx.x;
        const int64_t t = p < num_parts ? tiles_of(box_off[p + 1] - box_off[p], q_off[p + 1] - q_off[p]) : 0;
        s_scan[threadIdx.x] = t;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const int64_t v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (p < num_parts) tile_prefix[p] = s_carry + s_scan[threadIdx.x] - t;
        __syncthreads();
        if (threadIdx.x == 255) s_carry += s_scan[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_prefix[num_parts] = s_carry;
}

// One CTA = one 32-row x 256-column tile of one part's matrix: the thread keeps its column's constants in registers, the
// rows' constants are broadcast from shared memory, stores are coalesced along the columns.
template <int FL, int D3>
__global__ void __launch_bounds__(THREADS) kitti_pair_kernel(const Rec* __restrict__ rec_b, const Rec* __restrict__ rec_q,
                                                             const RecZ* __restrict__ z_b, const RecZ* __restrict__ z_q, Parts P,
                                                             int criterion, float* __restrict__ out) {
    __shared__ float4 s_row[TR];  // cx, cy, 1.01 r + 1e-4 (|cx| + |cy|), area
    __shared__ uint16_t s_queue[TR * TC];
    __shared__ int s_count;
    int64_t row0 = 0, col0 = 0, nrows = P.n, ncols = P.k, obase = 0, lt = blockIdx.x;
    if (P.box_off) {
        if ((int64_t)blockIdx.x >= P.tile_prefix[P.num_parts]) return;  // the grid is an upper bound
        int lo = 0, hi = P.num_parts - 1;
        while (lo < hi) {  // last part whose first tile is <= blockIdx.x; parts without tiles have equal prefixes and are skipped
            const int mid = (lo + hi + 1) >> 1;
            if (P.tile_prefix[mid] <= (int64_t)blockIdx.x) lo = mid;
            else hi = mid - 1;
        }
        row0 = P.box_off[lo], col0 = P.q_off[lo];
        nrows = P.box_off[lo + 1] - row0, ncols = P.q_off[lo + 1] - col0;
        obase = P.out_off[lo];
        lt = (int64_t)blockIdx.x - P.tile_prefix[lo];
    }
    const int64_t ctiles = (ncols + TC - 1) / TC;
    const int64_t tr = lt / ctiles, tc = lt - tr * ctiles;
    const int64_t r0 = tr * TR, c0 = tc * TC;
    const int nr = (int)min((int64_t)TR, nrows - r0);
    if (threadIdx.x == 0) s_count = 0;
    if (threadIdx.x < nr) {
        const float4 f = reinterpret_cast<const float4*>(rec_b + row0 + r0 + threadIdx.x)[2];
        s_row[threadIdx.x] = make_float4(f.x, f.y, f.z * 1.01f + 1e-4f * (fabsf(f.x) + fabsf(f.y)), f.w);
    }
    __syncthreads();
    const int64_t col = c0 + threadIdx.x;
    if (col < ncols) {
        const float4 fq = reinterpret_cast<const float4*>(rec_q + col0 + col)[2];
        const float aq = fq.z * 1.01f + 1e-4f * (fabsf(fq.x) + fabsf(fq.y)) + 1e-6f;
        float* o = out + obase + r0 * ncols + col;
#pragma unroll 4
        for (int r = 0; r < nr; ++r) {
            const float4 fb = s_row[r];
            // exact-zero test: centres further apart than the circum-radii plus a slack far above the rounding of the
            // reference's predicates (a few ulp of the coordinates) => < 3 polygon points => area exactly 0.0
            const float dx = fb.x - fq.x, dy = fb.y - fq.y;
            const float d2 = dx * dx + dy * dy;
            const float R = fb.z + aq;
            if (d2 > R * R && d2 < 3.0e38f) {
                o[(int64_t)r * ncols] = D3 ? 0.0f : zero_bev(fq.w, fb.w, criterion);
            } else {
                s_queue[atomicAdd(&s_count, 1)] = (uint16_t)((r << 8) | threadIdx.x);
            }
        }
    }
    __syncthreads();
    const int cnt = s_count;
    for (int t = threadIdx.x; t < cnt; t += THREADS) {
        const int q = s_queue[t], r = q >> 8, c = q & 255;
        const int64_t bi = row0 + r0 + r, qi = col0 + c0 + c;
        float cb[8], cq[8];
        const float4* pb = reinterpret_cast<const float4*>(rec_b + bi);
        const float4* pq = reinterpret_cast<const float4*>(rec_q + qi);
        const float4 b0 = pb[0], b1 = pb[1], b2 = pb[2], q0 = pq[0], q1 = pq[1], q2 = pq[2];
        cb[0] = b0.x, cb[1] = b0.y, cb[2] = b0.z, cb[3] = b0.w, cb[4] = b1.x, cb[5] = b1.y, cb[6] = b1.z, cb[7] = b1.w;
        cq[0] = q0.x, cq[1] = q0.y, cq[2] = q0.z, cq[3] = q0.w, cq[4] = q1.x, cq[5] = q1.y, cq[6] = q1.z, cq[7] = q1.w;
        const double ai = inter_area<FL>(cq, cb);  // the query box is the kernel's first argument (rotate_iou.py:289-291)
        float v;
        if (D3) v = finish_d3(__double2float_rn(ai), z_b[bi], z_q[qi], criterion);
        else v = finish_bev(ai, q2.w, b2.w, criterion);
        out[obase + (r0 + r) * ncols + c0 + c] = v;
    }
}

inline size_t ws_bytes(int64_t nb, int64_t nq, int parts) {
    return align_up((size_t)nb * sizeof(Rec), 256) + align_up((size_t)nq * sizeof(Rec), 256) + align_up((size_t)nb * sizeof(RecZ), 256) +
           align_up((size_t)nq * sizeof(RecZ), 256) + align_up((size_t)(parts + 1) * sizeof(int64_t), 256);
}

static int run(const float* b32, const float* q32, const double* b64, const double* q64, int64_t nb, int64_t nq, Parts P,
               int64_t total, int d3, int criterion, float* out, void* ws, size_t wsb, unsigned flags, void* stream) {
    const int np = P.box_off ? P.num_parts : 0;
    if (!ws || wsb < ws_bytes(nb, nq, np)) {
        set_error("workspace too small: need %zu bytes, got %zu", ws_bytes(nb, nq, np), wsb);
        return LG_ERR_WORKSPACE;
    }
    // tiles: exact for a single problem; for parts an upper bound from the totals (sum ceil(n/32) ceil(k/256) <=
    // total/8192 + rows/32 + cols/256 + parts), the CTAs beyond the real count return at once
    const int64_t tiles = P.box_off ? total / (TR * TC) + nb / TR + nq / TC + np + 3 : tiles_of(P.n, P.k);  // + 3: the three floors
    if (tiles > 2147483647LL) {
        set_error("%lld tiles exceed the grid limit; split the call", (long long)tiles);
        return LG_ERR_TOO_LARGE;
    }
    char* w = static_cast<char*>(ws);
    Rec* rb = reinterpret_cast<Rec*>(w);
    w += align_up((size_t)nb * sizeof(Rec), 256);
    Rec* rq = reinterpret_cast<Rec*>(w);
    w += align_up((size_t)nq * sizeof(Rec), 256);
    RecZ* zb = reinterpret_cast<RecZ*>(w);
    w += align_up((size_t)nb * sizeof(RecZ), 256);
    RecZ* zq = reinterpret_cast<RecZ*>(w);
    w += align_up((size_t)nq * sizeof(RecZ), 256);
    int64_t* tile_prefix = reinterpret_cast<int64_t*>(w);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool strict = flags & LG_FLAG_STRICT_FP32;
    const unsigned gb = (unsigned)((nb + 255) / 256), gq = (unsigned)((nq + 255) / 256);
    if (P.box_off) {
        kitti_tiles_kernel<<<1, 256, 0, st>>>(P.box_off, P.q_off, P.num_parts, tile_prefix);
        P.tile_prefix = tile_prefix;
    }
    if (strict) {
        kitti_prep_kernel<0><<<gb, 256, 0, st>>>(b32, b64, nb, rb, d3 ? zb : nullptr);
        kitti_prep_kernel<0><<<gq, 256, 0, st>>>(q32, q64, nq, rq, d3 ? zq : nullptr);
    } else {
        kitti_prep_kernel<1><<<gb, 256, 0, st>>>(b32, b64, nb, rb, d3 ? zb : nullptr);
        kitti_prep_kernel<1><<<gq, 256, 0, st>>>(q32, q64, nq, rq, d3 ? zq : nullptr);
    }
    int rc = check_launch("kitti_prep_kernel");
    if (rc) return rc;
    const unsigned grid = (unsigned)tiles;
    if (strict) {
        if (d3) kitti_pair_kernel<0, 1><<<grid, THREADS, 0, st>>>(rb, rq, zb, zq, P, criterion, out);
        else kitti_pair_kernel<0, 0><<<grid, THREADS, 0, st>>>(rb, rq, zb, zq, P, criterion, out);
    } else {
        if (d3) kitti_pair_kernel<1, 1><<<grid, THREADS, 0, st>>>(rb, rq, zb, zq, P, criterion, out);
        else kitti_pair_kernel<1, 0><<<grid, THREADS, 0, st>>>(rb, rq, zb, zq, P, criterion, out);
    }
    return check_launch("kitti_pair_kernel");
}

}  // namespace kitti
}  // namespace lg

extern "C" size_t lg_kitti_workspace_bytes(int64_t num_boxes, int64_t num_query_boxes, int num_parts) {
    if (num_boxes < 0 || num_query_boxes < 0 || num_parts < 0) return 0;
    return lg::kitti::ws_bytes(num_boxes, num_query_boxes, num_parts);
}

extern "C" int lg_rotate_iou_eval(const float* boxes, int64_t n, const float* query_boxes, int64_t k, float* out, int criterion,
                                  void* ws, size_t ws_bytes, unsigned flags, void* stream) {
    using namespace lg;
    if (n < 0 || k < 0) {
        set_error("negative size n=%lld k=%lld", (long long)n, (long long)k);
        return LG_ERR_INVALID_ARG;
    }
    if (n == 0 || k == 0) return LG_OK;
    if (!boxes || !query_boxes || !out) {
        set_error("null pointer (boxes=%p query_boxes=%p out=%p)", (const void*)boxes, (const void*)query_boxes, (void*)out);
        return LG_ERR_INVALID_ARG;
    }
    kitti::Parts P{nullptr, nullptr, nullptr, nullptr, 1, n, k};
    return kitti::run(boxes, query_boxes, nullptr, nullptr, n, k, P, n * k, 0, criterion, out, ws, ws_bytes, flags, stream);
}

extern "C" int lg_d3_box_overlap(const double* boxes, int64_t n, const double* qboxes, int64_t k, float* out, int criterion, void* ws,
                                 size_t ws_bytes, unsigned flags, void* stream) {
    using namespace lg;
    if (n < 0 || k < 0) {
        set_error("negative size n=%lld k=%lld", (long long)n, (long long)k);
        return LG_ERR_INVALID_ARG;
    }
    if (n == 0 || k == 0) return LG_OK;
    if (!boxes || !qboxes || !out) {
        set_error("null pointer (boxes=%p qboxes=%p out=%p)", (const void*)boxes, (const void*)qboxes, (void*)out);
        return LG_ERR_INVALID_ARG;
    }
    kitti::Parts P{nullptr, nullptr, nullptr, nullptr, 1, n, k};
    return kitti::run(nullptr, nullptr, boxes, qboxes, n, k, P, n * k, 1, criterion, out, ws, ws_bytes, flags, stream);
}

extern "C" int lg_kitti_overlaps_parts(const double* gt_boxes, int64_t num_gt, const double* dt_boxes, int64_t num_dt,
                                       const int64_t* gt_off, const int64_t* dt_off, const int64_t* out_off, int num_parts,
                                       int64_t num_out, int metric, int criterion, float* out, void* ws, size_t ws_bytes,
                                       unsigned flags, void* stream) {
    using namespace lg;
    if (num_gt < 0 || num_dt < 0 || num_parts < 0 || num_out < 0) {
        set_error("negative size (num_gt=%lld num_dt=%lld num_parts=%d num_out=%lld)", (long long)num_gt, (long long)num_dt, num_parts,
                  (long long)num_out);
        return LG_ERR_INVALID_ARG;
    }
    if (metric != 1 && metric != 2) {
        set_error("metric must be 1 (bev) or 2 (3d), got %d (metric 0, image boxes, has no rotated overlap)", metric);
        return LG_ERR_INVALID_ARG;
    }
    if (num_parts == 0 || num_out == 0) return LG_OK;
    if (!gt_boxes || !dt_boxes || !gt_off || !dt_off || !out_off || !out) {
        set_error("null pointer");
        return LG_ERR_INVALID_ARG;
    }
    kitti::Parts P{gt_off, dt_off, out_off, nullptr, num_parts, 0, 0};
    return kitti::run(nullptr, nullptr, gt_boxes, dt_boxes, num_gt, num_dt, P, num_out, metric == 2, criterion, out, ws, ws_bytes, flags, stream);
}
