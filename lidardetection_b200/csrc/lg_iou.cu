// lg_iou.cu -- N x M rotated BEV overlap / IoU / fused 3D IoU for sm_100a.
//
// Replaces boxes_overlap_kernel / boxes_iou_bev_kernel and their launchers
// (/root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265, 378-398) and fuses the ~12
// elementwise torch kernels of boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:48-81).
//
// Design (B200: 148 SMs, 228 KB smem/SM, HBM3e):
//   1. prep kernel: one thread per box builds a 112-byte record (lg_geom.cuh) -- all trigonometry,
//      corner rotation, edge vectors and margin arithmetic happens N+M times instead of N*M times -- and,
//      for the column boxes, a compact 16-byte cull quad (centre, radius) that the cull phase reads coalesced.
//   2. strip kernel (lg_strip.cuh): a CTA owns 64 rows (records in smem) x a run of columns, swept in
//      32 x 64 tiles (flat variant for M <= 64: 256 rows x M columns, linear pair index so that stores stay
//      coalesced for 20-column matrices).
//        cull   -- lanes run along columns; |ca - cb|^2 > (ra + rb)^2 proves the reference would
//                  return exactly +0.0, which is stored at once (coalesced st.global.cs, one full
//                  128-byte line per warp instruction); survivors are compacted into a smem queue;
//        drain  -- the queue keeps filling across the column tiles and is drained by ALL threads when
//                  it is nearly full, so the divergent polygon code runs with full warps whatever the
//                  survivor density (0.3 % for anchors x GT, 100 % for the dense microbench); column
//                  records are read through L1/L2 (they are 112 B x M, L2 resident); results are stored directly.
//      DRAM traffic == algorithmic bytes 4*N*M + 28*(N+M) (+ the 128*(N+M)-byte record round trip).
//      Sparse workloads are HBM-write bound, dense ones FP32-issue bound.  52 KB smem and <= 80 registers
//      per thread keep 3 CTAs (24 warps) of the dense build resident per SM.
//   3. large sparse matrices (>= 2^26 pairs): a two-phase sweep -- iou_sweep_kernel (cull, zero stores, survivor list) at the speed
//      of the store stream, then iou_pairs_kernel (the polygon path on the list); see the comment above iou_sweep_kernel.
//   4. 64-bit output offsets (the reference's int32 index overflows at 2^31 pairs).
#include "lg_common.cuh"
#include "lg_strip.cuh"

namespace lg {

// records of both box sets + the column cull quads + the density flag (16 B) + 16 B of slack, a multiple of 16 bytes
static inline size_t lg_iou_workspace_bytes_base(int64_t n, int64_t m) {
    return (size_t)(n + m) * REC_F4 * sizeof(float4) + (size_t)m * sizeof(float4) + 16 /* density flag */ + 16;
}

static inline unsigned pairs_grid() {  // iou_pairs_kernel: a grid-stride loop, three CTAs per SM
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static std::atomic<int> cache[64];
    if (dev >= 0 && dev < 64) {
        int v = cache[dev].load(std::memory_order_relaxed);
        if (v == 0) {
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cache[dev].store(sms, std::memory_order_relaxed);
        } else {
            sms = v;
        }
    }
    return (unsigned)(3 * sms);
}

template <int FL>
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ a, int64_t n, const float* __restrict__ b,
                                                   int64_t m, float4* __restrict__ rec_a, float4* __restrict__ rec_b,
                                                   float4* __restrict__ cull_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        make_record<FL>(a + i * 7, rec_a + i * REC_F4);
    } else if (i < n + m) {
        const int64_t j = i - n;
        make_record<FL>(b + j * 7, rec_b + j * REC_F4);
        cull_b[j] = rec_b[j * REC_F4 + REC_CULL];
    }
}

enum { MODE_OVERLAP = 0, MODE_IOU_BEV = 1, MODE_IOU3D = 2 };

__device__ __forceinline__ float finish_pair(const int mode, const float ov, const float4* A, const float4* B) {
    if (mode == MODE_IOU_BEV) return iou_from_overlap(ov, A[REC_CULL].w, B[REC_CULL].w);
    if (mode == MODE_IOU3D) return iou3d_from_overlap(ov, A[REC_Z], B[REC_Z]);
    return ov;
}

// ---- wide matrices: a CTA owns 64 rows x cols_per_cta columns --------------------------------------
constexpr int SK_ROWS = 64, SK_TROWS = 32, SK_TCOLS = 128;
constexpr int SK_SHIFT = 20;  // queue code = row << 20 | (column - first column of the CTA)
constexpr int SK_MAX_COLS = 1 << SK_SHIFT;

struct StripSmem {
    static constexpr size_t a_bytes = (size_t)SK_ROWS * REC_F4 * sizeof(float4);
    static constexpr size_t dup_bytes = (size_t)SK_ROWS * 2 * sizeof(float4);
    static constexpr size_t total = a_bytes + dup_bytes + ST_SLAB_BYTES_STAGED + (size_t)(ST_THREADS / 32) * (WQ_CAP + 1 + WR_CAP) * sizeof(uint32_t);
};

// A tile is 32 rows x 128 columns; a warp owns 8 of its rows (rsub + 4k) x 64 columns, two adjacent columns per
// lane.  The exact-zero cull of a full tile runs on packed FP32 pairs -- FADD2 / FMUL2 / FFMA2, the two columns of
// a lane in one instruction -- and stores the two zeros of a culled pair with one 8-byte st.global.cs.
// key of (value, index) for atomicMax: a larger value wins, equal values keep the LOWER index (torch.max's "first maximal value")
__device__ __forceinline__ unsigned long long max_key(float v, unsigned idx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

// REDUCE = false: write the N x M matrix.  REDUCE = true: write nothing of it; keep per-row / per-column (max, argmax)
// keys instead (what the target assigners and the evaluation actually consume) -- the 200k x 200k matrix
// (160 GB) is then never materialised.
// DENSE = true : polygon path inlined, 80 registers, 3 CTAs/SM -- best when a large share of the pairs overlap;
// DENSE = false: polygon path called out of line, 128 registers, 2 CTAs/SM, the sweep (cull + zero stores) stays in
//                registers -- best for the usual, sparse matrix.  Both are launched; `dense_flag` (written by
//                density_probe_kernel from a sample of the pairs, no host round trip) tells each whether it is its turn.
#ifndef LG_SK_DENSE_MINB
#define LG_SK_DENSE_MINB 3
#endif
#ifndef LG_SK_SPARSE_MINB
#define LG_SK_SPARSE_MINB 2
#endif
template <int FL, bool REDUCE, bool DENSE>
__device__ __forceinline__ void strip_body(const float4* __restrict__ rec_a, const int64_t n, const float4* __restrict__ rec_b,
                                           const float4* __restrict__ cull_b, const int64_t m, float* __restrict__ out, const int64_t ld,
                                           const int mode, const int cols_per_cta, const int64_t strips_m,
                                           unsigned long long* __restrict__ rowkey, unsigned long long* __restrict__ colkey,
                                           const int64_t strip) {
    constexpr int NT = ST_THREADS;
    extern __shared__ float4 smem4[];
    float4* sA = smem4;
    float4* sDup = sA + SK_ROWS * REC_F4;  // per row: (cx, cx, cy, cy), (rad, rad, -, -)
    float2* slab = reinterpret_cast<float2*>(sDup + SK_ROWS * 2);
    uint32_t* lists = reinterpret_cast<uint32_t*>(slab + SLAB_ROWS * NT);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t sn = strip / strips_m, sm = strip - sn * strips_m;
    const int64_t row0 = sn * SK_ROWS, col0 = sm * cols_per_cta;
    const int na = (int)min((int64_t)SK_ROWS, n - row0), nb = (int)min((int64_t)cols_per_cta, m - col0);
    const int ntiles = 2 * ((nb + SK_TCOLS - 1) / SK_TCOLS);  // (column tile, row half)

    for (int e = tid; e < na * REC_F4; e += NT) sA[e] = __ldg(rec_a + row0 * REC_F4 + e);
    if (tid < na) {
        const float4 c = __ldg(rec_a + (row0 + tid) * REC_F4 + REC_CULL);
        sDup[2 * tid] = make_float4(c.x, c.x, c.y, c.y);
        sDup[2 * tid + 1] = make_float4(c.z, c.z, 0.f, 0.f);
    }
    __syncthreads();  // the only CTA barrier: from here on the warps run independently

    const float4* const gB = rec_b + col0 * REC_F4;
    const float4* const gcull = cull_b + col0;
    float* const outb = out + row0 * ld + col0;
    // 8-byte stores need an even pitch and an 8-byte aligned first element of every row of this CTA
    const bool vec2_ok = REDUCE || (((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(outb) & 7) == 0));
    auto emit = [&](int r, int c, float ov, const float4* A, const float4* B) {
        const float v = finish_pair(mode, ov, A, B);
        if (REDUCE) {
            if (v > 0.f) {  // zeros never beat the initial key (0.0, index 0)
                if (rowkey) atomicMax(rowkey + row0 + r, max_key(v, (unsigned)col0 + (unsigned)c));
                if (colkey) atomicMax(colkey + col0 + c, max_key(v, (unsigned)row0 + (unsigned)r));
            }
        } else {
            __stcs(outb + (int64_t)r * ld + c, v);
        }
    };
    WarpQueue q;
    q.list = lists + warp * (WQ_CAP + 1 + WR_CAP);
    q.rare = q.list + WQ_CAP + 1;
    q.count = 0;
    q.rcount = 0;
    float2* const slab_warp = slab + warp * 32;

    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (!DENSE) {
        // ---- sparse build (the usual matrix: < 6 % of the pairs survive the cull): the sweep is a stream of zeros, so it is laid
        // out for the store -- FOUR adjacent columns per lane (one 16-byte st.global.cs per lane and row: a warp writes 512
        // contiguous bytes), four consecutive rows per warp and trip.  The zeros are stored UNCONDITIONALLY, straight after the
        // packed cull (no predicate, no branch); the few survivors go onto the warp's list and their results overwrite the zero
        // after the __syncwarp that closes the trip (same warp, ordered by the barrier).  ~13 instructions per 64 pairs.
        const int cl = 4 * lane;
        const unsigned lt = (1u << lane) - 1u;
        const bool vec4_ok = REDUCE || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(outb) & 15) == 0));
        float4 bn[4];
#pragma unroll
        for (int u = 0; u < 4; u++) bn[u] = cl + u < nb ? ld_keep(gcull + cl + u) : zero4;
        for (int t = 0;; t++) {  // one extra trip flushes the list (a single call site of the polygon path keeps the code small)
            const bool last = t >= ntiles;
            if (!last) {
                const int ct = (t >> 1) * SK_TCOLS, c = ct + cl;  // this lane's first column, relative to col0
                float4 b[4];
#pragma unroll
                for (int u = 0; u < 4; u++) b[u] = bn[u];
                if (t & 1) {  // both row halves of this column tile use b; fetch the next tile's quads now
                    const int cn = c + SK_TCOLS;
#pragma unroll
                    for (int u = 0; u < 4; u++) bn[u] = cn + u < nb ? ld_keep(gcull + cn + u) : zero4;
                }
                const int rbase = (t & 1) * SK_TROWS + warp * 4;  // this warp's four rows of the trip
                if (rbase < na) {
                    float* outp = outb + (int64_t)rbase * ld + c;
                    auto push4 = [&](const int row, const bool s0, const bool s1, const bool s2, const bool s3) {
                        if (!__any_sync(0xffffffffu, s0 | s1 | s2 | s3)) return;  // the usual case: one vote per row
                        const bool sv[4] = {s0, s1, s2, s3};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const unsigned mu = __ballot_sync(0xffffffffu, sv[u]);
                            if (sv[u]) q.list[q.count + __popc(mu & lt)] = (uint32_t)((row << SK_SHIFT) | (c + u));
                            q.count += __popc(mu);
                        }
                    };
                    if (vec4_ok && rbase + 4 <= na && ct + SK_TCOLS <= nb) {  // full rows x full column tile: packed cull, no bounds tests
                        const f32x2 bx01 = pack2(b[0].x, b[1].x), by01 = pack2(b[0].y, b[1].y), br01 = pack2(b[0].z, b[1].z);
                        const f32x2 bx23 = pack2(b[2].x, b[3].x), by23 = pack2(b[2].y, b[3].y), br23 = pack2(b[2].z, b[3].z);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const ulonglong2 axy = *reinterpret_cast<const ulonglong2*>(sDup + 2 * (rbase + k));  // (cx,cx), (cy,cy)
                            const f32x2 ar = *reinterpret_cast<const f32x2*>(sDup + 2 * (rbase + k) + 1);         // (rad,rad)
                            const f32x2 dxa = sub2(axy.x, bx01), dya = sub2(axy.y, by01), rra = add2(ar, br01);
                            const f32x2 dxb = sub2(axy.x, bx23), dyb = sub2(axy.y, by23), rrb = add2(ar, br23);
                            const f32x2 d2a = fma2(dxa, dxa, mul2(dya, dya)), r2a = mul2(rra, rra);
                            const f32x2 d2b = fma2(dxb, dxb, mul2(dyb, dyb)), r2b = mul2(rrb, rrb);
                            float d0, d1, d2, d3, r0, r1, r2, r3;
                            unpack2(d2a, d0, d1);
                            unpack2(r2a, r0, r1);
                            unpack2(d2b, d2, d3);
                            unpack2(r2b, r2, r3);
                            if (!REDUCE) {
                                __stcs(reinterpret_cast<float4*>(outp), zero4);  // culled or not: exactly +0.0 first
                                outp += ld;
                            }
                            push4(rbase + k, !(d0 > r0), !(d1 > r1), !(d2 > r2), !(d3 > r3));  // NaN => keep: the polygon path decides
                        }
                    } else {
#pragma unroll 1
                        for (int k = 0; k < 4; k++) {
                            const int r = rbase + k;
                            bool sv[4] = {false, false, false, false};
                            if (r < na) {
                                const float4 ac = sA[r * REC_F4 + REC_CULL];
#pragma unroll
                                for (int u = 0; u < 4; u++)
                                    if (c + u < nb) {
                                        sv[u] = cull_survives(ac, b[u]);
                                        if (!REDUCE && !sv[u]) __stcs(outp + u, 0.f);
                                    }
                            }
                            outp += ld;
                            push4(r, sv[0], sv[1], sv[2], sv[3]);
                        }
                    }
                }
                __syncwarp();
            }
            while (q.count >= (last ? 1 : 32)) warp_round<FL, SK_SHIFT, false>(q, sA, gB, slab_warp, NT, lane, emit);
            if (last) break;
        }
        if (q.rcount > 0) warp_drain_rare<FL, SK_SHIFT>(q, sA, gB, slab_warp, NT, lane, emit);
        return;
    }
    const int rsub = warp >> 1, cbase = (warp & 1) * 64 + 2 * lane;
    float4 bn0 = cbase < nb ? __ldg(gcull + cbase) : zero4, bn1 = cbase + 1 < nb ? __ldg(gcull + cbase + 1) : zero4;
    for (int t = 0;; t++) {  // one extra trip flushes the list (a single call site of the polygon path keeps the code small)
        const bool last = t >= ntiles;
        if (!last) {
            const int c = (t >> 1) * SK_TCOLS + cbase;  // first of this lane's two columns, relative to col0
            const float4 b0 = bn0, b1 = bn1;
            if (t & 1) {  // both row halves of this column tile use b0, b1; fetch the next tile's quads now
                const int cn = c + SK_TCOLS;
                bn0 = cn < nb ? __ldg(gcull + cn) : zero4;
                bn1 = cn + 1 < nb ? __ldg(gcull + cn + 1) : zero4;
            }
            const int rhalf = (t & 1) * SK_TROWS;
            if (rhalf >= na) continue;
            const int rbase = rhalf + rsub;
            float* outp = outb + (int64_t)rbase * ld + c;
            const int64_t ostep = 4 * ld;
            // survivors go straight onto the warp's list (a warp-uniform, rarely taken branch when the matrix is sparse)
            const unsigned lt = (1u << lane) - 1u;
            auto push2 = [&](const int row, const bool s0, const bool s1) {
                if (!__any_sync(0xffffffffu, s0 | s1)) return;  // the usual case of a sparse matrix: one vote per row
                const unsigned m0 = __ballot_sync(0xffffffffu, s0), m1 = __ballot_sync(0xffffffffu, s1);
                {
                    if (s0) q.list[q.count + __popc(m0 & lt)] = (uint32_t)((row << SK_SHIFT) | c);
                    q.count += __popc(m0);
                    if (s1) q.list[q.count + __popc(m1 & lt)] = (uint32_t)((row << SK_SHIFT) | (c + 1));
                    q.count += __popc(m1);
                }
            };
            if (vec2_ok && rhalf + SK_TROWS <= na && (t >> 1) * SK_TCOLS + SK_TCOLS <= nb) {  // full tile: packed cull, no bounds tests
                const f32x2 bx = pack2(b0.x, b1.x), by = pack2(b0.y, b1.y), br = pack2(b0.z, b1.z);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const ulonglong2 axy = *reinterpret_cast<const ulonglong2*>(sDup + 2 * (rbase + 4 * k));  // (cx,cx), (cy,cy)
                    const f32x2 ar = *reinterpret_cast<const f32x2*>(sDup + 2 * (rbase + 4 * k) + 1);         // (rad,rad)
                    const f32x2 dx = sub2(axy.x, bx), dy = sub2(axy.y, by), rr = add2(ar, br);
                    const f32x2 d2 = fma2(dx, dx, mul2(dy, dy)), r2 = mul2(rr, rr);
                    float d2a, d2b, r2a, r2b;
                    unpack2(d2, d2a, d2b);
                    unpack2(r2, r2a, r2b);
                    const bool s0 = !(d2a > r2a), s1 = !(d2b > r2b);  // NaN => keep: the polygon path decides
                    if (!REDUCE) {
                        if (!s0 && !s1) {
                            __stcs(reinterpret_cast<float2*>(outp), make_float2(0.f, 0.f));  // culled: exactly +0.0, written once
                        } else {
                            if (!s0) __stcs(outp, 0.f);
                            if (!s1) __stcs(outp + 1, 0.f);
                        }
                        outp += ostep;
                    }
                    push2(rbase + 4 * k, s0, s1);
                }
            } else {
                const bool v0 = c < nb, v1 = c + 1 < nb;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int r = rbase + 4 * k;
                    bool s0 = false, s1 = false;
                    if (r < na) {
                        const float4 ac = sA[r * REC_F4 + REC_CULL];
                        if (v0) {
                            s0 = cull_survives(ac, b0);
                            if (!REDUCE && !s0) __stcs(outp, 0.f);
                        }
                        if (v1) {
                            s1 = cull_survives(ac, b1);
                            if (!REDUCE && !s1) __stcs(outp + 1, 0.f);
                        }
                    }
                    outp += ostep;
                    push2(r, s0, s1);
                }
            }
            __syncwarp();
        }
        while (q.count >= (last ? 1 : 32)) warp_round<FL, SK_SHIFT, DENSE>(q, sA, gB, slab_warp, NT, lane, emit);
        if (last) break;
    }
    if (q.rcount > 0) warp_drain_rare<FL, SK_SHIFT>(q, sA, gB, slab_warp, NT, lane, emit);
}

// cta_only == nullptr: one strip per CTA.  Otherwise (the two-phase path's complete sweep): a small persistent grid walks the
// strips and redoes only those whose flag is raised -- normally none, so the launch must cost next to nothing.
template <int FL, bool REDUCE, bool DENSE>
__global__ void __launch_bounds__(ST_THREADS, DENSE ? LG_SK_DENSE_MINB : LG_SK_SPARSE_MINB)
    iou_strip_kernel(const float4* __restrict__ rec_a, const int64_t n, const float4* __restrict__ rec_b,
                     const float4* __restrict__ cull_b, const int64_t m, float* __restrict__ out, const int64_t ld,
                     const int mode, const int cols_per_cta, const int64_t strips_m, unsigned long long* __restrict__ rowkey,
                     unsigned long long* __restrict__ colkey, const int* __restrict__ dense_flag, const int* __restrict__ cta_only,
                     const int64_t total_strips) {
    if (dense_flag && (*dense_flag != 0) != DENSE) return;
    if (!cta_only) {
        strip_body<FL, REDUCE, DENSE>(rec_a, n, rec_b, cull_b, m, out, ld, mode, cols_per_cta, strips_m, rowkey, colkey, blockIdx.x);
        return;
    }
    for (int64_t s = blockIdx.x; s < total_strips; s += gridDim.x) {
        if (cta_only[s] == 0) continue;  // CTA-uniform
        strip_body<FL, REDUCE, DENSE>(rec_a, n, rec_b, cull_b, m, out, ld, mode, cols_per_cta, strips_m, rowkey, colkey, s);
        __syncthreads();  // the next strip re-uses the shared memory
    }
}

// ---------------------------------------------------------------------------------------------------
// Two-phase form of the sparse sweep (large matrices): the matrix is a stream of zeros with a few survivors, so the sweep
// and the polygon path are separate kernels with the resources each needs.
//   iou_sweep_kernel   cull + unconditional 16-byte zero stores + the survivors' (row, column) pairs appended to ONE global list;
//                      no records, no polygon code: few registers, a full SM of warps -- it runs at the speed of the store stream;
//   iou_pairs_kernel   the polygon path on the list, one pair per lane, always on full warps (the dense build's inlined path);
//                      results overwrite the zeros (a later launch of the same stream).
// The list has a fixed capacity (SweepPlan).  A strip whose survivors do not fit raises its flag in cta_over and stops; the
// complete one-kernel sweep (iou_strip_kernel<., ., false>) then redoes exactly the flagged strips -- any density is handled,
// the usual one (< 1 % survivors) never gets there.  list entry: row << 32 | column (global indices); ~0 = hole left by a
// reservation that straddled the capacity.
constexpr int SW_LIST = 31 + 512;        // per-warp staging: leftover + one trip (4 rows x 128 columns)
constexpr int SW_FLUSH = 128;            // staged survivors that trigger a flush (one global atomicAdd per flush)
constexpr unsigned long long SW_HOLE = ~0ull;
struct SweepCtrl {
    unsigned long long count;  // list entries reserved so far (may exceed the capacity)
    unsigned long long pad;
};

template <bool REDUCE>
__global__ void __launch_bounds__(ST_THREADS)
    iou_sweep_kernel(const float4* __restrict__ rec_a, const int64_t n, const float4* __restrict__ cull_b, const int64_t m,
                     float* __restrict__ out, const int64_t ld, const int cols_per_cta, const int64_t strips_m,
                     const int* __restrict__ dense_flag, SweepCtrl* __restrict__ ctrl, int* __restrict__ cta_over,
                     unsigned long long* __restrict__ list, const unsigned long long cap) {
    if (dense_flag && *dense_flag != 0) return;
    __shared__ float4 sDup[SK_ROWS * 2];  // per row: (cx, cx, cy, cy), (rad, rad, -, -)
    __shared__ uint32_t lists[(ST_THREADS / 32) * SW_LIST];
    __shared__ int s_over;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t strip = blockIdx.x;
    const int64_t sn = strip / strips_m, sm = strip - sn * strips_m;
    const int64_t row0 = sn * SK_ROWS, col0 = sm * cols_per_cta;
    const int na = (int)min((int64_t)SK_ROWS, n - row0), nb = (int)min((int64_t)cols_per_cta, m - col0);
    const int ntc = (nb + SK_TCOLS - 1) / SK_TCOLS;  // column tiles
    if (tid < SK_ROWS) {
        const float4 c = tid < na ? __ldg(rec_a + (row0 + tid) * REC_F4 + REC_CULL) : make_float4(0.f, 0.f, 0.f, 0.f);
        sDup[2 * tid] = make_float4(c.x, c.x, c.y, c.y);
        sDup[2 * tid + 1] = make_float4(c.z, c.z, 0.f, 0.f);
    }
    if (tid == 0) s_over = 0;
    __syncthreads();  // the only CTA barrier
    const float4* const gcull = cull_b + col0;
    float* const outb = REDUCE ? nullptr : out + row0 * ld + col0;
    const bool vec4_ok = REDUCE || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(outb) & 15) == 0));
    uint32_t* const wl = lists + warp * SW_LIST;
    int count = 0;  // warp-uniform
    const unsigned lt = (1u << lane) - 1u;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cl = 4 * lane;
    auto flush = [&]() {  // the staged survivors go to the global list: one reservation per flush
        __syncwarp();
        unsigned long long base = 0ull;
        if (lane == 0) base = atomicAdd(&ctrl->count, (unsigned long long)count);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + (unsigned long long)count <= cap) {
            for (int i = lane; i < count; i += 32) {
                const uint32_t e = wl[i];
                list[base + i] = ((unsigned long long)(row0 + (e >> SK_SHIFT)) << 32) | (unsigned long long)(col0 + (e & (SK_MAX_COLS - 1)));
            }
        } else {
            for (unsigned long long i = base + lane; i < cap; i += 32) list[i] = SW_HOLE;  // the part of the reservation below the capacity
            if (lane == 0) {
                cta_over[blockIdx.x] = 1;
                s_over = 1;
            }
        }
        count = 0;
        __syncwarp();
    };
    float4 bn[4];
#pragma unroll
    for (int u = 0; u < 4; u++) bn[u] = cl + u < nb ? ld_keep(gcull + cl + u) : zero4;
    for (int tc = 0; tc < ntc; tc++) {
        if (*(volatile int*)&s_over) break;  // this strip goes to the complete sweep: stop producing
        const int ct = tc * SK_TCOLS, c = ct + cl;  // this lane's first column, relative to col0
        float4 b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) b[u] = bn[u];
        {
            const int cn = c + SK_TCOLS;  // the next tile's quads: in flight during this tile's 8 rows
#pragma unroll
            for (int u = 0; u < 4; u++) bn[u] = cn + u < nb ? ld_keep(gcull + cn + u) : zero4;
        }
        const f32x2 bx01 = pack2(b[0].x, b[1].x), by01 = pack2(b[0].y, b[1].y), br01 = pack2(b[0].z, b[1].z);
        const f32x2 bx23 = pack2(b[2].x, b[3].x), by23 = pack2(b[2].y, b[3].y), br23 = pack2(b[2].z, b[3].z);
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const int rbase = half * SK_TROWS + warp * 4;  // this warp's four rows of the half
            if (rbase >= na) break;
            auto push4 = [&](const int row, const bool s0, const bool s1, const bool s2, const bool s3) {
                if (!__any_sync(0xffffffffu, s0 | s1 | s2 | s3)) return;  // the usual case: one vote per row
                const bool sv[4] = {s0, s1, s2, s3};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned mu = __ballot_sync(0xffffffffu, sv[u]);
                    if (sv[u]) wl[count + __popc(mu & lt)] = (uint32_t)((row << SK_SHIFT) | (c + u));
                    count += __popc(mu);
                }
            };
            if (vec4_ok && rbase + 4 <= na && ct + SK_TCOLS <= nb) {  // full rows x full column tile: packed cull, no bounds tests
                float* outp = REDUCE ? nullptr : outb + (int64_t)rbase * ld + c;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const ulonglong2 axy = *reinterpret_cast<const ulonglong2*>(sDup + 2 * (rbase + k));  // (cx,cx), (cy,cy)
                    const f32x2 ar = *reinterpret_cast<const f32x2*>(sDup + 2 * (rbase + k) + 1);         // (rad,rad)
                    const f32x2 dxa = sub2(axy.x, bx01), dya = sub2(axy.y, by01), rra = add2(ar, br01);
                    const f32x2 dxb = sub2(axy.x, bx23), dyb = sub2(axy.y, by23), rrb = add2(ar, br23);
                    const f32x2 d2a = fma2(dxa, dxa, mul2(dya, dya)), r2a = mul2(rra, rra);
                    const f32x2 d2b = fma2(dxb, dxb, mul2(dyb, dyb)), r2b = mul2(rrb, rrb);
                    float d0, d1, d2, d3, r0, r1, r2, r3;
                    unpack2(d2a, d0, d1);
                    unpack2(r2a, r0, r1);
                    unpack2(d2b, d2, d3);
                    unpack2(r2b, r2, r3);
                    if (!REDUCE) {
                        __stcs(reinterpret_cast<float4*>(outp), zero4);  // survivor or not: exactly +0.0 first
                        outp += ld;
                    }
                    push4(rbase + k, !(d0 > r0), !(d1 > r1), !(d2 > r2), !(d3 > r3));  // NaN => keep: the polygon path decides
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 4; k++) {
                    const int r = rbase + k;
                    bool sv[4] = {false, false, false, false};
                    if (r < na) {
                        const float4 ac = make_float4(sDup[2 * r].x, sDup[2 * r].z, sDup[2 * r + 1].x, 0.f);
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (c + u < nb) {
                                sv[u] = cull_survives(ac, b[u]);
                                if (!REDUCE) __stcs(outb + (int64_t)r * ld + c + u, 0.f);
                            }
                    }
                    push4(r, sv[0], sv[1], sv[2], sv[3]);
                }
            }
            if (count >= SW_FLUSH) flush();
        }
    }
    if (count > 0) flush();
}

// The polygon path on the survivor list: a grid-stride loop, 32 consecutive entries per warp and trip, in two stages.  Most
// survivors of the circle test are near misses whose polygon is empty: stage 1 (phase A: the 16 straddle + 8 margin tests) settles
// them.  The few pairs that do overlap (3..8 vertices) are queued -- pair + masks, 16 bytes -- and stage 2 (crossing points, order,
// fan) runs whenever the warp's queue holds 32 of them, so the divergent code sees full warps instead of the one or two overlapping
// pairs of a trip.
template <int FL, bool REDUCE>
__global__ void __launch_bounds__(ST_THREADS, 3)
    iou_pairs_kernel(const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, const unsigned long long* __restrict__ list,
                     const SweepCtrl* __restrict__ ctrl, const unsigned long long cap, float* __restrict__ out, const int64_t ld,
                     const int mode, unsigned long long* __restrict__ rowkey, unsigned long long* __restrict__ colkey,
                     const int* __restrict__ dense_flag) {
    if (dense_flag && *dense_flag != 0) return;
    constexpr int NT = ST_THREADS, NW = NT / 32, RARE = 64, HEAVY = 64;
    __shared__ float2 slab[SLAB_ROWS_SMEM_B * NT];
    __shared__ unsigned long long rare[NW * RARE];
    __shared__ ulonglong2 heavyq[NW * HEAVY];  // (pair, xmask | cmask << 16)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long total = min(ctrl->count, cap);
    float2* const slab_warp = slab + warp * 32;
    unsigned long long* const wr = rare + warp * RARE;
    ulonglong2* const wh = heavyq + warp * HEAVY;
    int rcount = 0, hcount = 0;  // warp-uniform
    const unsigned lt = (1u << lane) - 1u;
    auto emit = [&](const unsigned long long e, const float ov, const float4* A, const float4* B) {
        const float v = finish_pair(mode, ov, A, B);
        const unsigned r = (unsigned)(e >> 32), c = (unsigned)e;
        if (REDUCE) {
            if (v > 0.f) {  // zeros never beat the initial key (0.0, index 0)
                if (rowkey) atomicMax(rowkey + r, max_key(v, c));
                if (colkey) atomicMax(colkey + c, max_key(v, r));
            }
        } else {
            __stcs(out + (int64_t)r * ld + c, v);
        }
    };
    // the deferred pairs of this warp (> 8 vertices, angular near-ties): lanes 0..7, 16 vertex + 16 angle slots each
    auto drain_rare_list = [&]() {
        __syncwarp();
        if (lane < 8) {
            auto slab16 = [&](int k) -> float2& { return slab_warp[(k & 7) * NT + lane + ((k >> 3) << 3)]; };
            auto ang16 = [&](int k) -> float& { return reinterpret_cast<float*>(slab_warp + (k >> 1) * NT + 16 + lane)[k & 1]; };
            for (int i = lane; i < rcount; i += 8) {
                const unsigned long long e = wr[i];
                const float4* A = rec_a + (int64_t)(e >> 32) * REC_F4;
                const float4* B = rec_b + (int64_t)(e & 0xFFFFFFFFull) * REC_F4;
                float ov = overlap_area16<FL>(A, B, slab16);
                if (ov < 0.f) ov = overlap_area_slow<FL>(A, B, slab16, ang16);
                emit(e, ov, A, B);
            }
        }
        rcount = 0;
        __syncwarp();
    };
    auto defer_rare = [&](const bool defer, const unsigned long long e) {  // all lanes
        const unsigned dm = __ballot_sync(0xffffffffu, defer);
        if (dm) {
            if (defer) wr[rcount + __popc(dm & lt)] = e;
            rcount += __popc(dm);
            if (rcount > RARE - 32) drain_rare_list();
        }
    };
    // stage 2 on the top n (<= 32) entries of the warp's queue
    auto heavy_round = [&](const int n) {
        __syncwarp();
        const bool act = lane < n;
        const unsigned hm = __ballot_sync(0xffffffffu, act);
        bool defer = false;
        unsigned long long e = 0ull;
        if (act) {
            const ulonglong2 h = wh[hcount - n + lane];
            e = h.x;
            const uint32_t xm = (uint32_t)h.y & 0xFFFFu, cm = ((uint32_t)h.y >> 16) & 0xFFu;
            const float4* A = rec_a + (int64_t)(e >> 32) * REC_F4;
            const float4* B = rec_b + (int64_t)(e & 0xFFFFFFFFull) * REC_F4;
            const float ov = overlap_area_heavy<FL, false>(A, B, xm, cm, __popc(xm) + __popc(cm), slab_warp + lane, NT, hm);
            if (ov < 0.f) defer = true;  // angular near-tie
            else emit(e, ov, A, B);
        }
        hcount -= n;
        defer_rare(defer, e);
        __syncwarp();
    };
    const unsigned long long stride = (unsigned long long)gridDim.x * NW * 32;
    // the kernel is bound by the latency of its gathers (list entry -> two 112-byte records somewhere in the L2): the entry of the
    // NEXT trip is loaded one trip ahead and its records are requested into the L1 while this trip's pairs are worked on
    auto fetch = [&](const unsigned long long idx) -> unsigned long long {
        unsigned long long e = SW_HOLE;
        if (idx < total) e = __ldg(list + idx);
        if (e != SW_HOLE) {
            const char* pa = reinterpret_cast<const char*>(rec_a + (int64_t)(e >> 32) * REC_F4);
            const char* pb = reinterpret_cast<const char*>(rec_b + (int64_t)(e & 0xFFFFFFFFull) * REC_F4);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pa));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + 96));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pb));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + 96));
        }
        return e;
    };
    const unsigned long long base0 = ((unsigned long long)blockIdx.x * NW + warp) * 32;
    unsigned long long e_next = fetch(base0 + lane);
    for (unsigned long long base = base0; base < total; base += stride) {
        const unsigned long long e = e_next;
        e_next = fetch(base + stride + lane);
        const bool act = e != SW_HOLE;
        bool defer = false, heavy = false;
        uint32_t xm = 0u, cm = 0u;
        if (act) {
            const float4* A = rec_a + (int64_t)(e >> 32) * REC_F4;
            const float4* B = rec_b + (int64_t)(e & 0xFFFFFFFFull) * REC_F4;
            pair_masks<FL, false>(A, B, xm, cm);
            const int cnt = __popc(xm) + __popc(cm);
            if (cnt <= 2) emit(e, 0.f, A, B);  // the fan sum is empty or a single zero term
            else if (cnt > 8) defer = true;
            else heavy = true;
        }
        const unsigned hv = __ballot_sync(0xffffffffu, heavy);
        if (hv) {
            if (heavy) wh[hcount + __popc(hv & lt)] = make_ulonglong2(e, (unsigned long long)(xm | (cm << 16)));
            hcount += __popc(hv);
        }
        defer_rare(defer, e);
        if (hcount >= 32) heavy_round(32);
    }
    if (hcount > 0) heavy_round(hcount);
    if (rcount > 0) drain_rare_list();
}

// Survivor density of the exact-zero cull, estimated from 8192 pseudo-random pairs: flag = 1 when more than ~6 % survive.
__global__ void __launch_bounds__(1024) density_probe_kernel(const float4* __restrict__ rec_a, const int64_t n, const float4* __restrict__ cull_b,
                                                              const int64_t m, int* __restrict__ flag) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    // 8 samples per thread, all 16 loads in flight at once (the probe is a chain of cache misses, nothing else)
    float4 qa[8], qb[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int s = threadIdx.x + 1024 * u;
        const unsigned long long h = (unsigned long long)(s + 1) * 0x9E3779B97F4A7C15ull;
        const int64_t i = (int64_t)((h >> 33) % (unsigned long long)n), j = (int64_t)(((h * 0xD1B54A32D192ED03ull) >> 33) % (unsigned long long)m);
        qa[u] = __ldg(rec_a + i * REC_F4 + REC_CULL);
        qb[u] = __ldg(cull_b + j);
    }
    int mine = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) mine += cull_survives(qa[u], qb[u]) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) *flag = cnt > 512 ? 1 : 0;
}

// ---- narrow matrices (M <= 64, e.g. anchors x GT): a CTA owns 256 rows x all M columns, flat pair index ----
constexpr int FLAT_ROWS = 256;
constexpr int FLAT_COLS = 64;
constexpr int FLAT_UNROLL = 8;

struct FlatSmem {
    static constexpr size_t a_bytes = (size_t)FLAT_ROWS * REC_F4 * sizeof(float4);
    static constexpr size_t b_bytes = (size_t)FLAT_COLS * REC_F4 * sizeof(float4);
    static constexpr size_t total = a_bytes + b_bytes + ST_SLAB_BYTES + (size_t)(ST_QCAP + ST_RARECAP) * sizeof(uint16_t);
};

template <int FL>
__global__ void __launch_bounds__(ST_THREADS, 3)
    iou_flat_kernel(const float* __restrict__ box_a, const int64_t n, const float* __restrict__ box_b, const int m,
                    float* __restrict__ out, const int64_t ld, const int mode, const unsigned inv_m /* ceil(2^20 / m) */) {
    constexpr int NT = ST_THREADS;
    extern __shared__ float4 smem4[];
    float4* sA = smem4;
    float4* sB = sA + FLAT_ROWS * REC_F4;
    float2* slab = reinterpret_cast<float2*>(sB + FLAT_COLS * REC_F4);
    uint16_t* queue = reinterpret_cast<uint16_t*>(slab + SLAB_ROWS_SMEM_B * NT);
    uint16_t* rareq = queue + ST_QCAP;
    __shared__ int qcount, rcount;

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * FLAT_ROWS;
    const int na = (int)min((int64_t)FLAT_ROWS, n - row0);
    // the records are built in place (no prep kernel, no 112-byte round trip through HBM: for a 20-column matrix that
    // round trip would be three times the matrix itself); the M <= 64 column records are rebuilt by every CTA
    if (tid < na) make_record<FL>(box_a + (row0 + tid) * 7, sA + tid * REC_F4);
    if (tid < m) make_record<FL>(box_b + (int64_t)tid * 7, sB + tid * REC_F4);
    if (tid == 0) {
        qcount = 0;
        rcount = 0;
    }
    float* const outb = out + row0 * ld;
    auto emit = [&](int r, int c, float ov, const float4* A, const float4* B) {
        __stcs(outb + (int64_t)r * ld + c, finish_pair(mode, ov, A, B));
    };

    const int npairs = na * m;  // <= 16384
    for (int chunk = 0;; chunk += FLAT_UNROLL * NT) {
        __syncthreads();
        const int qn = qcount, rn = rcount;
        __syncthreads();
        const bool last = chunk >= npairs;
        if (last || qn > ST_QCAP - FLAT_UNROLL * NT) {
            if (rn + qn > ST_RARECAP) {
                drain_rare<FL, 6>(sA, sB, slab, rareq, &rcount, emit);
                __syncthreads();
            }
            drain_main<FL, 6>(sA, sB, slab, queue, qn, rareq, &rcount, emit);
            if (last) {
                __syncthreads();
                drain_rare<FL, 6>(sA, sB, slab, rareq, &rcount, emit);
                break;
            }
            if (tid == 0) qcount = 0;
            __syncthreads();
        }
        unsigned mk[FLAT_UNROLL];
        int codes[FLAT_UNROLL];
#pragma unroll
        for (int k = 0; k < FLAT_UNROLL; k++) {
            const int e = chunk + k * NT + tid;  // consecutive lanes -> consecutive pairs -> consecutive addresses
            bool surv = false;
            codes[k] = 0;
            if (e < npairs) {
                const int r = (int)(((unsigned)e * inv_m) >> 20), c = e - r * m;  // exact for e < 2^14, m <= 64
                codes[k] = (r << 6) | c;
                surv = cull_survives(sA[r * REC_F4 + REC_CULL], sB[c * REC_F4 + REC_CULL]);
                if (!surv) __stcs(outb + (int64_t)r * ld + c, 0.f);
            }
            mk[k] = __ballot_sync(0xffffffffu, surv);
        }
        unsigned any = 0u;
#pragma unroll
        for (int k = 0; k < FLAT_UNROLL; k++) any |= mk[k];
        if (any) {
            int total = 0;
#pragma unroll
            for (int k = 0; k < FLAT_UNROLL; k++) total += __popc(mk[k]);
            int base = 0;
            if (lane == 0) base = atomicAdd(&qcount, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < FLAT_UNROLL; k++) {
                if (mk[k]) {
                    if ((mk[k] >> lane) & 1u) queue[base + __popc(mk[k] & lt)] = (uint16_t)codes[k];
                    base += __popc(mk[k]);
                }
            }
        }
    }
}

// strip decomposition of an N x M sweep and, for large matrices, the two-phase plan (survivor list in the workspace)
struct SweepPlan {
    int64_t cols, strips_m, strips;
    bool two_phase;
    unsigned long long cap;  // list entries
    size_t extra_bytes;      // control block + per-strip flags + list, appended to the workspace
    size_t flags_bytes;
};
static SweepPlan sweep_plan(int64_t n, int64_t m, unsigned flags) {
    SweepPlan p = {};
    if (n <= 0 || m <= 0) return p;
    // columns per CTA: as long a run as still leaves >= ~16 CTAs per SM-slot for load balance (148 SMs x 3 CTAs)
    const int64_t strips_n = (n + SK_ROWS - 1) / SK_ROWS;
    int64_t want_m = (16 * 444 + strips_n - 1) / strips_n;  // column splits wanted
    if (want_m < 1) want_m = 1;
    int64_t cols = (m + want_m - 1) / want_m;
    cols = (cols + SK_TCOLS - 1) / SK_TCOLS * SK_TCOLS;
    if (cols < 2 * SK_TCOLS) cols = 2 * SK_TCOLS;
    if (cols > SK_MAX_COLS) cols = SK_MAX_COLS;
    p.cols = cols;
    p.strips_m = (m + cols - 1) / cols;
    p.strips = strips_n * p.strips_m;
    // two-phase from 2^26 pairs on (below that the extra launches cost more than the sweep saves: measured, tools/time_two_phase.py); indices are packed in 32 bits
    const double pairs = (double)n * (double)m;
    p.two_phase = !(flags & LG_FLAG_IOU_ONE_KERNEL) && m > FLAT_COLS && pairs >= 67108864.0 && n < 0xFFFFFFFFLL && m < 0xFFFFFFFFLL && p.strips <= 0x7fffffffLL;
    if (p.two_phase) {
        double cap = pairs / 32.0;  // ~3 % survivors; the usual matrix has < 1 %
        if (cap < 65536.0) cap = 65536.0;
        if (cap > 33554432.0) cap = 33554432.0;  // 256 MB
        p.cap = (unsigned long long)cap;
        p.flags_bytes = align_up((size_t)p.strips * sizeof(int), 16);
        p.extra_bytes = sizeof(SweepCtrl) + p.flags_bytes + (size_t)p.cap * sizeof(unsigned long long);
        if (flags & LG_FLAG_IOU_SMALL_LIST) p.cap = 1024;  // (testing) same layout, tiny capacity: the overflow path
    }
    return p;
}

static int check_args(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws, size_t ws_bytes) {
    if (n < 0 || m < 0) {
        set_error("negative size n=%lld m=%lld", (long long)n, (long long)m);
        return LG_ERR_INVALID_ARG;
    }
    if (n == 0 || m == 0) return LG_OK;
    if (!a || !b || !out) {
        set_error("null pointer (boxes_a=%p boxes_b=%p out=%p)", (const void*)a, (const void*)b, (void*)out);
        return LG_ERR_INVALID_ARG;
    }
    if (ld < m) {
        set_error("ld_out=%lld < m=%lld", (long long)ld, (long long)m);
        return LG_ERR_INVALID_ARG;
    }
    if (!ws || ws_bytes < lg_iou_workspace_bytes(n, m) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("workspace %p of %zu B; need %zu B, 16-byte aligned", ws, ws_bytes, lg_iou_workspace_bytes(n, m));
        return LG_ERR_WORKSPACE;
    }
    return LG_OK;
}

template <int FL>
static int run_iou(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws, int mode,
                   cudaStream_t st, unsigned flags) {
    int rc;
    if (m <= FLAT_COLS) {
        const int64_t ctas = (n + FLAT_ROWS - 1) / FLAT_ROWS;
        if (ctas > 0x7fffffffLL) {
            set_error("%lld row blocks exceed the 1-D grid limit; split the call by row blocks", (long long)ctas);
            return LG_ERR_TOO_LARGE;
        }
        auto kern = iou_flat_kernel<FL>;
        if ((rc = set_smem(kern, FlatSmem::total))) return rc;
        const unsigned inv_m = (unsigned)(((1u << 20) + (unsigned)m - 1u) / (unsigned)m);
        kern<<<(unsigned)ctas, ST_THREADS, FlatSmem::total, st>>>(a, n, b, (int)m, out, ld, mode, inv_m);
        return check_launch("iou_flat_kernel");
    }
    float4* ra = reinterpret_cast<float4*>(ws);
    float4* rb = ra + n * REC_F4;
    float4* cb = rb + m * REC_F4;
    const int64_t total = n + m;
    prep_kernel<FL><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, n, b, m, ra, rb, cb);
    if ((rc = check_launch("prep_kernel"))) return rc;
    const SweepPlan pl = sweep_plan(n, m, flags);
    const int64_t cols = pl.cols, strips_m = pl.strips_m, strips = pl.strips;
    if (strips > 0x7fffffffLL) {
        set_error("%lld strips exceed the 1-D grid limit; split the call by row blocks", (long long)strips);
        return LG_ERR_TOO_LARGE;
    }
    int* flag = reinterpret_cast<int*>(cb + m);
    density_probe_kernel<<<1, 1024, 0, st>>>(ra, n, cb, m, flag);
    if ((rc = check_launch("density_probe_kernel"))) return rc;
    auto ks = iou_strip_kernel<FL, false, false>;
    auto kd = iou_strip_kernel<FL, false, true>;
    if ((rc = set_smem(ks, StripSmem::total))) return rc;
    if ((rc = set_smem(kd, StripSmem::total))) return rc;
    const int* only = nullptr;
    if (pl.two_phase) {
        // sparse matrices: zeros + survivor list, then the polygon path on the list; the complete sweep below redoes only the strips
        // whose survivors did not fit (none, normally)
        char* extra = reinterpret_cast<char*>(ws) + lg_iou_workspace_bytes_base(n, m);
        SweepCtrl* ctrl = reinterpret_cast<SweepCtrl*>(extra);
        int* over = reinterpret_cast<int*>(extra + sizeof(SweepCtrl));
        unsigned long long* list = reinterpret_cast<unsigned long long*>(extra + sizeof(SweepCtrl) + pl.flags_bytes);
        cudaError_t e = cudaMemsetAsync(extra, 0, sizeof(SweepCtrl) + pl.flags_bytes, st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
            return (int)e;
        }
        iou_sweep_kernel<false><<<(unsigned)strips, ST_THREADS, 0, st>>>(ra, n, cb, m, out, ld, (int)cols, strips_m, flag, ctrl, over, list, pl.cap);
        if ((rc = check_launch("iou_sweep_kernel"))) return rc;
        iou_pairs_kernel<FL, false><<<pairs_grid(), ST_THREADS, 0, st>>>(ra, rb, list, ctrl, pl.cap, out, ld, mode, nullptr, nullptr, flag);
        if ((rc = check_launch("iou_pairs_kernel"))) return rc;
        only = over;
    }
    const unsigned gs = only ? (unsigned)min(strips, (int64_t)pairs_grid()) : (unsigned)strips;  // persistent when it only mops up
    ks<<<gs, ST_THREADS, StripSmem::total, st>>>(ra, n, rb, cb, m, out, ld, mode, (int)cols, strips_m, nullptr, nullptr, flag, only, strips);
    kd<<<(unsigned)strips, ST_THREADS, StripSmem::total, st>>>(ra, n, rb, cb, m, out, ld, mode, (int)cols, strips_m, nullptr, nullptr, flag, nullptr, strips);
    return check_launch("iou_strip_kernel");
}

__global__ void __launch_bounds__(256) key_init_kernel(unsigned long long* __restrict__ k, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) k[i] = 0x00000000FFFFFFFFull;  // (value 0.0, index 0)
}

__global__ void __launch_bounds__(256) key_unpack_kernel(const unsigned long long* __restrict__ k, int64_t count, float* __restrict__ vmax,
                                                         int64_t* __restrict__ arg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const unsigned long long key = k[i];
        if (vmax) vmax[i] = __uint_as_float((unsigned)(key >> 32));
        if (arg) arg[i] = (int64_t)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    }
}

template <int FL>
static int run_iou_reduce(const float* a, int64_t n, const float* b, int64_t m, void* ws, int mode, float* row_max, int64_t* row_arg,
                          float* col_max, int64_t* col_arg, cudaStream_t st, unsigned flags) {
    float4* ra = reinterpret_cast<float4*>(ws);
    float4* rb = ra + n * REC_F4;
    float4* cb = rb + m * REC_F4;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(cb + m + 1);  // (the 16 bytes after cb are the matrix path's density flag)
    const bool want_rows = row_max || row_arg, want_cols = col_max || col_arg;
    unsigned long long* rowkey = want_rows ? keys : nullptr;
    unsigned long long* colkey = want_cols ? keys + n : nullptr;
    int rc;
    prep_kernel<FL><<<(unsigned)((n + m + 255) / 256), 256, 0, st>>>(a, n, b, m, ra, rb, cb);
    if ((rc = check_launch("prep_kernel"))) return rc;
    key_init_kernel<<<(unsigned)((n + m + 255) / 256), 256, 0, st>>>(keys, n + m);
    if ((rc = check_launch("key_init_kernel"))) return rc;
    const SweepPlan pl = sweep_plan(n, m, flags);
    const int64_t cols = pl.cols, strips_m = pl.strips_m, strips = pl.strips;
    if (strips > 0x7fffffffLL) {
        set_error("%lld strips exceed the 1-D grid limit; split the call by row blocks", (long long)strips);
        return LG_ERR_TOO_LARGE;
    }
    auto kern = iou_strip_kernel<FL, true, false>;
    if ((rc = set_smem(kern, StripSmem::total))) return rc;
    const int* only = nullptr;
    if (pl.two_phase) {  // see run_iou; there is no density probe here: a dense matrix overflows the list at once and every strip falls back
        char* extra = reinterpret_cast<char*>(ws) + lg_iou_workspace_bytes_base(n, m) + (size_t)(n + m) * sizeof(unsigned long long);
        SweepCtrl* ctrl = reinterpret_cast<SweepCtrl*>(extra);
        int* over = reinterpret_cast<int*>(extra + sizeof(SweepCtrl));
        unsigned long long* list = reinterpret_cast<unsigned long long*>(extra + sizeof(SweepCtrl) + pl.flags_bytes);
        cudaError_t e = cudaMemsetAsync(extra, 0, sizeof(SweepCtrl) + pl.flags_bytes, st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
            return (int)e;
        }
        iou_sweep_kernel<true><<<(unsigned)strips, ST_THREADS, 0, st>>>(ra, n, cb, m, nullptr, 0, (int)cols, strips_m, nullptr, ctrl, over, list, pl.cap);
        if ((rc = check_launch("iou_sweep_kernel<reduce>"))) return rc;
        iou_pairs_kernel<FL, true><<<pairs_grid(), ST_THREADS, 0, st>>>(ra, rb, list, ctrl, pl.cap, nullptr, 0, mode, rowkey, colkey, nullptr);
        if ((rc = check_launch("iou_pairs_kernel<reduce>"))) return rc;
        only = over;
    }
    const unsigned gs = only ? (unsigned)min(strips, (int64_t)pairs_grid()) : (unsigned)strips;
    kern<<<gs, ST_THREADS, StripSmem::total, st>>>(ra, n, rb, cb, m, nullptr, 0, mode, (int)cols, strips_m, rowkey, colkey, nullptr, only, strips);
    if ((rc = check_launch("iou_strip_kernel<reduce>"))) return rc;
    if (want_rows) key_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowkey, n, row_max, row_arg);
    if (want_cols) key_unpack_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(colkey, m, col_max, col_arg);
    return check_launch("key_unpack_kernel");
}

static int iou_entry(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws, size_t ws_bytes,
                     unsigned flags, void* stream, int mode) {
    int rc = check_args(a, n, b, m, out, ld, ws, ws_bytes);
    if (rc || n == 0 || m == 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (flags & LG_FLAG_STRICT_FP32) return run_iou<0>(a, n, b, m, out, ld, ws, mode, st, flags);
    return run_iou<1>(a, n, b, m, out, ld, ws, mode, st, flags);
}

}  // namespace lg

extern "C" size_t lg_iou_workspace_bytes(int64_t n, int64_t m) {
    if (n < 0 || m < 0) return 0;
    return lg::lg_iou_workspace_bytes_base(n, m) + lg::sweep_plan(n, m, 0).extra_bytes;
}

extern "C" size_t lg_iou_reduce_workspace_bytes(int64_t n, int64_t m) {
    if (n < 0 || m < 0) return 0;
    return lg::lg_iou_workspace_bytes_base(n, m) + (size_t)(n + m) * sizeof(unsigned long long) + lg::sweep_plan(n, m, 0).extra_bytes;
}

extern "C" int lg_boxes_iou_reduce(const float* a, int64_t n, const float* b, int64_t m, int kind, float* row_max, int64_t* row_argmax,
                                   float* col_max, int64_t* col_argmax, void* ws, size_t ws_bytes, unsigned flags, void* stream) {
    using namespace lg;
    if (n < 0 || m < 0 || kind < 0 || kind > 2) {
        set_error("bad argument n=%lld m=%lld kind=%d", (long long)n, (long long)m, kind);
        return LG_ERR_INVALID_ARG;
    }
    if (n >= 0xFFFFFFFFLL || m >= 0xFFFFFFFFLL) {
        set_error("n=%lld m=%lld: indices are packed in 32 bits", (long long)n, (long long)m);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n == 0 || m == 0) {  // empty reductions: (0, 0) for whatever side has entries, as a matrix of no columns has no maximum
        if (n > 0 && row_max) cudaMemsetAsync(row_max, 0, sizeof(float) * n, st);
        if (n > 0 && row_argmax) cudaMemsetAsync(row_argmax, 0, sizeof(int64_t) * n, st);
        if (m > 0 && col_max) cudaMemsetAsync(col_max, 0, sizeof(float) * m, st);
        if (m > 0 && col_argmax) cudaMemsetAsync(col_argmax, 0, sizeof(int64_t) * m, st);
        return check_launch("cudaMemsetAsync");
    }
    if (!a || !b) {
        set_error("null pointer (boxes_a=%p boxes_b=%p)", (const void*)a, (const void*)b);
        return LG_ERR_INVALID_ARG;
    }
    if (!ws || ws_bytes < lg_iou_reduce_workspace_bytes(n, m) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("workspace %p of %zu B; need %zu B, 16-byte aligned", ws, ws_bytes, lg_iou_reduce_workspace_bytes(n, m));
        return LG_ERR_WORKSPACE;
    }
    if (flags & LG_FLAG_STRICT_FP32) return run_iou_reduce<0>(a, n, b, m, ws, kind, row_max, row_argmax, col_max, col_argmax, st, flags);
    return run_iou_reduce<1>(a, n, b, m, ws, kind, row_max, row_argmax, col_max, col_argmax, st, flags);
}

extern "C" int lg_boxes_overlap_bev(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws,
                                    size_t ws_bytes, unsigned flags, void* stream) {
    return lg::iou_entry(a, n, b, m, out, ld, ws, ws_bytes, flags, stream, lg::MODE_OVERLAP);
}

extern "C" int lg_boxes_iou_bev(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws,
                                size_t ws_bytes, unsigned flags, void* stream) {
    return lg::iou_entry(a, n, b, m, out, ld, ws, ws_bytes, flags, stream, lg::MODE_IOU_BEV);
}

extern "C" int lg_boxes_iou3d(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws,
                              size_t ws_bytes, unsigned flags, void* stream) {
    return lg::iou_entry(a, n, b, m, out, ld, ws, ws_bytes, flags, stream, lg::MODE_IOU3D);
}
