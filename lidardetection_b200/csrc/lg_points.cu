// lg_points.cu -- points-in-boxes for sm_100a.
//
// Replaces points_in_boxes_kernel + launcher (/root/reference/pcdet/ops/roiaware_pool3d/src/
// roiaware_pool3d_kernel.cu:16-36, 313-359) and offers the all-pairs mask form of points_in_boxes_cpu
// (roiaware_pool3d.cpp:121-168) on the GPU.  The predicate and its bit-exactness argument are in lg_pib.cuh.
//
// Index form (B frames x M points x T boxes -> first containing box or -1): the reference does M*T predicate
// evaluations with per-pair trigonometry; here a CTA
//   1. builds the T 32-byte box records (trigonometry hoisted) and their conservative BEV footprints,
//   2. lays a uniform grid over the frame's boxes in shared memory -- per cell one 32-bit list of up to four
//      candidate box indices in ascending order (16 KB: 4096 cells) -- all threads sharing the flattened
//      (box, cell) pairs: separating-axis test, sorted insert by compare-and-swap,
//   3. streams its points through shared memory with 1-D bulk async copies (TMA, cp.async.bulk + mbarrier,
//      two 12 KB stages per CTA, three CTAs per SM -- the sizes were swept on the B200: occupancy matters more here than
//      grid resolution or pipeline depth); points whose cell lists a candidate go to per-warp work lists
//      and are tested on full warps against the (<= 4) boxes of their cell in ascending index order.
// HBM traffic is the algorithmic 16 B per point (+ 28 T per CTA); the kernel is HBM-bound when the batch is
// large enough to fill the machine.  Frames whose boxes have a non-finite footprint fall back to
// testing every box.
#include "lg_common.cuh"
#include "lg_pib.cuh"

namespace lg {

constexpr int PIB_THREADS = 256;
#ifndef LG_PIB_TILE
#define LG_PIB_TILE 1024
#endif
constexpr int PIB_TILE = LG_PIB_TILE;          // points per stage
#ifndef LG_PIB_STAGES
#define LG_PIB_STAGES 2
#endif
#ifndef LG_PIB_CELLS
#define LG_PIB_CELLS 4096
#endif
#ifndef LG_PIB_MINB
#define LG_PIB_MINB 4
#endif
constexpr int PIB_STAGES = LG_PIB_STAGES;
constexpr int PIB_TILE_BYTES = PIB_TILE * 12;  // 12 KB
constexpr int PIB_CELLS = LG_PIB_CELLS;         // 16 KB of 32-bit candidate lists (lg_pib.cuh)
constexpr float PIB_MIN_CELL = 0.6f;            // cells no smaller than 0.6 x the mean footprint half-extent
constexpr int PIB_WARPS = PIB_THREADS / 32;
constexpr int PIB_WPTS = PIB_TILE / PIB_WARPS;  // 128 points of a tile per warp
constexpr int PIB_WLIST = 64;                   // per-warp work list (float4 items): <= 31 left over + <= 32 appended per slot of the step

// ---- mbarrier / bulk-copy wrappers (PTX ISA: cp.async.bulk, mbarrier) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct PibSmem {
    // dynamic smem layout: [stages][cells][work lists][records][box cell ranges][prefix]
    static constexpr size_t stage_bytes = (size_t)PIB_STAGES * PIB_TILE_BYTES;
    static constexpr size_t cell_bytes = (size_t)PIB_CELLS * sizeof(uint32_t);
    static constexpr size_t list_bytes = (size_t)PIB_WARPS * PIB_WLIST * sizeof(float4);
    // the scratch of the grid build (cell ranges, touch constants, prefix) is dead once the lists are built: it shares the
    // work lists' region
    __host__ __device__ static size_t build_bytes(int T) {
        const int tc = T < PIB_COMPACT_MAX_BOXES ? T : PIB_COMPACT_MAX_BOXES;
        return (size_t)tc * (sizeof(int4) + 2 * sizeof(float4)) + (size_t)(tc + 8) * sizeof(int);
    }
    __host__ __device__ static size_t region_bytes(int T) {
        const size_t b = (build_bytes(T) + 15) / 16 * 16;
        return b > list_bytes ? b : list_bytes;
    }
    static size_t total(int T) { return stage_bytes + cell_bytes + region_bytes(T) + (size_t)T * 2 * sizeof(float4); }
};

template <int FL>
__global__ void __launch_bounds__(PIB_THREADS, LG_PIB_MINB)
    pib_grid_kernel(const float* __restrict__ boxes, const float* __restrict__ pts, int32_t* __restrict__ out, const int T,
                    const int64_t M, const int64_t pts_per_cta) {
    constexpr int NT = PIB_THREADS;
    extern __shared__ float4 smem4[];
    float* stage = reinterpret_cast<float*>(smem4);
    uint32_t* cells = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem4) + PibSmem::stage_bytes);
    float4* lists = reinterpret_cast<float4*>(cells + PIB_CELLS);
    float4* srec = reinterpret_cast<float4*>(reinterpret_cast<char*>(lists) + PibSmem::region_bytes(T));
    int4* srange = reinterpret_cast<int4*>(lists);  // per box: ix0, iy0, cells per row, cells (build scratch, aliases the lists)
    float4* stouch = reinterpret_cast<float4*>(srange + min(T, PIB_COMPACT_MAX_BOXES));  // per box: pib_touch_consts
    int* sprefix = reinterpret_cast<int*>(stouch + 2 * min(T, PIB_COMPACT_MAX_BOXES));
    __shared__ uint64_t bars[PIB_STAGES];
    __shared__ float red[8][6];
    __shared__ PibGrid sgrid;
    __shared__ int s_use_grid, s_nvalid, s_total;
    __shared__ int s_done[PIB_STAGES];  // warps that have finished with the stage's current tile

    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p_begin = (int64_t)blockIdx.x * pts_per_cta;
    const int64_t p_end = min(M, p_begin + pts_per_cta);
    const int64_t npts = p_end - p_begin;
    const float* gp = pts + ((int64_t)b * M + p_begin) * 3;
    int32_t* go = out + (int64_t)b * M + p_begin;
    const int ntiles = (int)((npts + PIB_TILE - 1) / PIB_TILE);
    const bool tma_ok = (reinterpret_cast<uintptr_t>(gp) & 15) == 0;  // bulk copies need 16-byte aligned sources

    auto tile_bytes = [&](int t) -> unsigned { return (unsigned)(min((int64_t)PIB_TILE, npts - (int64_t)t * PIB_TILE) * 12); };
    auto issue = [&](int t) {  // thread 0 only; a tile whose byte count is not a multiple of 16 is loaded by the fallback
        const unsigned bytes = tile_bytes(t);
        if (tma_ok && (bytes & 15) == 0) {
            uint64_t* bar = &bars[t % PIB_STAGES];
            mbar_expect_tx(bar, bytes);
            bulk_g2s(stage + (size_t)(t % PIB_STAGES) * (PIB_TILE * 3), gp + (int64_t)t * PIB_TILE * 3, bytes, bar);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < PIB_STAGES; s++) s_done[s] = 0;
        for (int s = 0; s < PIB_STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int t = 0; t < min(ntiles, PIB_STAGES); t++) issue(t);  // the points start flowing while the grid is built
    }


    // ---- 1. records, footprints, frame bounds
    const float* fb = boxes + (int64_t)b * T * 7;
    float lo_x = INFINITY, hi_x = -INFINITY, lo_y = INFINITY, hi_y = -INFINITY, sum_ext = 0.f, nvalid = 0.f;
    bool bounded = true;
    for (int k = tid; k < T; k += NT) {
        float4 r0, r1;
        make_pib_record<FL>(fb + k * 7, 1e-5f, r0, r1);
        srec[2 * k] = r0;
        srec[2 * k + 1] = r1;
        float ex, ey;
        if (pib_footprint(r0, r1, ex, ey, bounded)) {
            lo_x = fminf(lo_x, r0.x - ex);
            hi_x = fmaxf(hi_x, r0.x + ex);
            lo_y = fminf(lo_y, r0.y - ey);
            hi_y = fmaxf(hi_y, r0.y + ey);
            sum_ext += 0.5f * (ex + ey);
            nvalid += 1.f;
        }
    }
    const unsigned all_bounded = __ballot_sync(0xffffffffu, bounded) == 0xffffffffu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, d));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, d));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, d));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, d));
        sum_ext += __shfl_xor_sync(0xffffffffu, sum_ext, d);
        nvalid += __shfl_xor_sync(0xffffffffu, nvalid, d);
    }
    if (lane == 0) {
        red[warp][0] = lo_x;
        red[warp][1] = hi_x;
        red[warp][2] = lo_y;
        red[warp][3] = hi_y;
        red[warp][4] = sum_ext;
        red[warp][5] = all_bounded ? nvalid : -1e30f;  // poison: any unbounded box disables the grid
    }
    {
        const uint4 e4 = make_uint4(PIB_CELL_EMPTY, PIB_CELL_EMPTY, PIB_CELL_EMPTY, PIB_CELL_EMPTY);
        for (int w = tid; w < PIB_CELLS / 4; w += NT) reinterpret_cast<uint4*>(cells)[w] = e4;
    }
    __syncthreads();
    if (tid == 0) {
        float a = INFINITY, bb = -INFINITY, c = INFINITY, d = -INFINITY, se = 0.f, nv = 0.f;
        for (int w = 0; w < NT / 32; w++) {
            a = fminf(a, red[w][0]);
            bb = fmaxf(bb, red[w][1]);
            c = fminf(c, red[w][2]);
            d = fmaxf(d, red[w][3]);
            se += red[w][4];
            nv += red[w][5];
        }
        const bool ok = nv >= 0.f && T <= PIB_COMPACT_MAX_BOXES;  // more boxes than the 8-bit lists can name: every box is tested
        s_use_grid = ok ? 1 : 0;
        s_nvalid = nv > 0.f ? 1 : 0;
        if (ok && nv > 0.f) sgrid = pib_make_grid(a, bb, c, d, se / nv, PIB_CELLS, PIB_MIN_CELL);
    }
    __syncthreads();
    const bool use_grid = s_use_grid != 0, any_valid = s_nvalid != 0;
    const PibGrid g = sgrid;

    // ---- 2. candidate lists.  (a) one thread per box: the cell range of its footprint; (b) prefix sum of the cell
    //         counts; (c) ALL threads share the flattened (box, cell) pairs: separating-axis test, sorted insert (CAS)
    if (use_grid && any_valid) {
        for (int k = tid; k < T; k += NT) {
            const float4 r0 = srec[2 * k], r1 = srec[2 * k + 1];
            float ex, ey;
            bool dummy = true;
            int4 rg = make_int4(0, 0, 1, 0);
            if (pib_footprint(r0, r1, ex, ey, dummy)) {
                const int ix0 = pib_cell_clamped(r0.x - ex, g.x0, g.invx, g.nx), ix1 = pib_cell_clamped(r0.x + ex, g.x0, g.invx, g.nx);
                const int iy0 = pib_cell_clamped(r0.y - ey, g.y0, g.invy, g.ny), iy1 = pib_cell_clamped(r0.y + ey, g.y0, g.invy, g.ny);
                rg = make_int4(ix0, iy0, ix1 - ix0 + 1, (ix1 - ix0 + 1) * (iy1 - iy0 + 1));
                float4 m0, m1;
                pib_touch_consts(r0, r1, g, m0, m1);
                stouch[2 * k] = m0;
                stouch[2 * k + 1] = m1;
            }
            srange[k] = rg;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive prefix over T <= 254 counts: 8 per lane
            int loc[8], sum = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = lane * 8 + u;
                loc[u] = sum;
                sum += k < T ? srange[k].w : 0;
            }
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const int excl = incl - sum;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = lane * 8 + u;
                if (k <= T) sprefix[k] = excl + loc[u];
            }
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        const int total = s_total;
        for (int idx = tid; idx < total; idx += NT) {
            int lo = 0, hi = T;  // the box whose pair range [prefix[k], prefix[k+1]) holds idx
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sprefix[mid] <= idx) lo = mid;
                else hi = mid;
            }
            const int4 rg = srange[lo];
            const int local = idx - sprefix[lo];
            const int cy = local / rg.z, cx = local - cy * rg.z;
            const int ix = rg.x + cx, iy = rg.y + cy;
            if (pib_cell_touches_fast(stouch[2 * lo], stouch[2 * lo + 1], ix, iy)) {
                uint32_t* cw = cells + iy * g.nx + ix;
                uint32_t old = *cw, assumed;
                do {
                    assumed = old;
                    old = atomicCAS(cw, assumed, pib_compact_insert(assumed, (uint32_t)lo));
                } while (old != assumed);
            }
        }
    }
    __syncthreads();

    // ---- 3. the points, a tile per trip.  Each warp owns 128 consecutive points of the tile.
    //   step     one point per lane (conflict-free LDS): cell lookup; points whose cell lists a candidate box
    //            are appended -- coordinates, cell and output offset -- to the warp's work list;
    //   fill     the tile's results are pre-set to -1 with one 16-byte store per lane;
    //   rounds   whenever the list holds 32 items, all 32 lanes test one item each against its cell's (<= 4)
    //            boxes in ascending order and overwrite the -1 of a point that is inside one (first hit = lowest box).
    // The list carries over from tile to tile, so the predicate always runs on full warps.
    float4* wlist = lists + warp * PIB_WLIST;
    int wcount = 0;
    auto round = [&](const int n) {  // the top n (<= 32) items of the list
        const int idx = wcount - n + lane;
        const bool act = lane < n;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t ids = PIB_CELL_EMPTY;
        if (act) {
            e = wlist[idx];
            ids = cells[__float_as_uint(e.w) >> 16];
        }
        int found = -1;
#pragma unroll
        for (int sl = 0; sl < 4; sl++) {
            const uint32_t id = (ids >> (8 * sl)) & 0xffu;
            const bool go_on = found < 0 && id < PIB_ID_MORE;
            if (__any_sync(0xffffffffu, go_on)) {
                if (go_on && pt_in_box<FL>(e.x, e.y, e.z, srec[2 * id], srec[2 * id + 1])) found = (int)id;
            }
        }
        if (found < 0 && (ids >> 24) == PIB_ID_MORE) {  // rare: a cell with more than four candidates
            for (int k = (int)((ids >> 16) & 0xffu) + 1; k < T; k++)
                if (pt_in_box<FL>(e.x, e.y, e.z, srec[2 * k], srec[2 * k + 1])) {
                    found = k;
                    break;
                }
        }
        if (found >= 0) go[__float_as_uint(e.w) & 0xffffu] = found;
        wcount -= n;
        __syncwarp();
    };
    // one 128-point slice: wp = the slice in shared memory, wnp = valid points in it, obase = offset of its first point
    // within the CTA's chunk (< 65536)
    auto process = [&](const float* wp, const int wnp, const int obase) {
        if (use_grid) {
            const int i0 = lane * 4;
            // fill: -1 for the warp's 128 points (results of hits are written over it after the __syncwarp below)
            {
                int32_t* o = go + obase + i0;
                if (i0 + 4 <= wnp && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                    __stcs(reinterpret_cast<int4*>(o), make_int4(-1, -1, -1, -1));
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i0 + u < wnp) o[u] = -1;
                }
            }
            // four consecutive points per lane (the same four its 16-byte fill covers): three conflict-free 16-byte
            // shared loads instead of twelve scalar ones, and four independent cell lookups in flight
            const float4* q4 = reinterpret_cast<const float4*>(wp + (size_t)i0 * 3);
            const float4 qa = q4[0], qb = q4[1], qc = q4[2];
            const float xs[4] = {qa.x, qa.w, qb.z, qc.y}, ys[4] = {qa.y, qb.x, qb.w, qc.z}, zs[4] = {qa.z, qb.y, qc.x, qc.w};
            int cell[4];
            bool has[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int ix = __float2int_rd(pib_cellf(xs[u], g.x0, g.invx)), iy = __float2int_rd(pib_cellf(ys[u], g.y0, g.invy));
                has[u] = any_valid && i0 + u < wnp && (unsigned)ix < (unsigned)g.nx && (unsigned)iy < (unsigned)g.ny;
                cell[u] = has[u] ? iy * g.nx + ix : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) has[u] = has[u] && cells[cell[u]] != PIB_CELL_EMPTY;
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned hm = __ballot_sync(0xffffffffu, has[u]);
                if (has[u])
                    wlist[wcount + __popc(hm & lt)] =
                        make_float4(xs[u], ys[u], zs[u], __uint_as_float(((uint32_t)cell[u] << 16) | (uint32_t)(obase + i0 + u)));
                wcount += __popc(hm);
                if (wcount >= 32) {  // keeps the list within 31 + 32 entries
                    __syncwarp();    // orders the fill and the list writes before the round
                    round(32);
                }
            }
        } else {
            // general path (more than 254 boxes, or a frame with an unbounded box): every box is tested, as the reference does
            for (int st = 0; st < 4; st++) {
                const int p = st * 32 + lane;
                if (p >= wnp) continue;
                const float* q = wp + p * 3;
                const float x = q[0], y = q[1], z = q[2];
                int r = -1;
                for (int k = 0; k < T; k++)
                    if (pt_in_box<FL>(x, y, z, srec[2 * k], srec[2 * k + 1])) {
                        r = k;
                        break;
                    }
                go[obase + p] = r;
            }
        }
    };
    for (int t = 0; t < ntiles; t++) {
        const int s = t % PIB_STAGES;
        float* sp = stage + (size_t)s * (PIB_TILE * 3);
        const unsigned bytes = tile_bytes(t);
        const int np = (int)(bytes / 12);
        if (tma_ok && (bytes & 15) == 0) {
            mbar_wait(&bars[s], (unsigned)((t / PIB_STAGES) & 1));
        } else {
            __syncthreads();  // (rare path) every warp is done with the tile this stage held
            const float* src = gp + (int64_t)t * PIB_TILE * 3;
            for (int i = tid; i < np * 3; i += NT) sp[i] = __ldg(src + i);
            __syncthreads();
        }
        const int wbase = warp * PIB_WPTS;       // this warp's first point within the tile
        process(sp + wbase * 3, np - wbase, t * PIB_TILE + wbase);
        __syncwarp();  // orders the fill above, and the list writes, before the rounds; the warp has read its points
        // No CTA barrier per tile: a warp that is done with the stage signs off, and the LAST one to do so refills it (tile
        // t + PIB_STAGES) -- nobody waits for the slowest warp except through the data it needs next.  The work-list items carry
        // their coordinates, so the rounds below no longer touch the stage.
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&s_done[s], 1) == PIB_WARPS - 1) {
                s_done[s] = 0;  // the next sign-off for this stage comes after the refill issued here has landed
                __threadfence_block();
                if (t + PIB_STAGES < ntiles) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(t + PIB_STAGES);
                }
            }
        }
        if (use_grid)
            while (wcount >= 32) round(32);
    }
    if (use_grid && wcount > 0) round(wcount);
}

constexpr int PIBM_BOXES = 32;  // boxes per CTA in the mask form

template <int FL>
__global__ void __launch_bounds__(PIB_THREADS)
    pib_mask_kernel(const float* __restrict__ boxes, const int64_t n, const float* __restrict__ pts, const int64_t m,
                    int32_t* __restrict__ out, const float margin) {
    __shared__ float4 srec[2 * PIBM_BOXES];
    const int tid = threadIdx.x;
    const int64_t box0 = (int64_t)blockIdx.y * PIBM_BOXES;
    const int nbx = (int)min((int64_t)PIBM_BOXES, n - box0);
    if (tid < nbx) {
        float4 r0, r1;
        make_pib_record<FL>(boxes + (box0 + tid) * 7, margin, r0, r1);
        srec[2 * tid] = r0;
        srec[2 * tid + 1] = r1;
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * PIB_THREADS + tid;
    if (p >= m) return;
    const float x = __ldg(pts + p * 3), y = __ldg(pts + p * 3 + 1), z = __ldg(pts + p * 3 + 2);
    for (int k = 0; k < nbx; k++) {
        const int v = pt_in_box<FL>(x, y, z, srec[2 * k], srec[2 * k + 1]) ? 1 : 0;
        __stcs(out + (box0 + k) * m + p, v);
    }
}

}  // namespace lg

extern "C" size_t lg_points_in_boxes_workspace_bytes(int, int, int64_t) { return 0; }

extern "C" int lg_points_in_boxes(const float* boxes, const float* pts, int32_t* out, int B, int T, int64_t M, void*, size_t,
                                  unsigned flags, void* stream) {
    using namespace lg;
    if (B < 0 || T < 0 || M < 0) {
        set_error("negative size batch=%d boxes=%d points=%lld", B, T, (long long)M);
        return LG_ERR_INVALID_ARG;
    }
    if (B == 0 || M == 0) return LG_OK;
    if (!pts || !out || (T > 0 && !boxes)) {
        set_error("null pointer (boxes=%p pts=%p out=%p)", (const void*)boxes, (const void*)pts, (void*)out);
        return LG_ERR_INVALID_ARG;
    }
    if (T > LG_PIB_MAX_BOXES) {
        set_error("num_boxes=%d exceeds LG_PIB_MAX_BOXES=%d", T, LG_PIB_MAX_BOXES);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // points per CTA: a whole frame when the batch alone fills the machine (the grid is built once per CTA),
    // otherwise split the frame, but never below 4 tiles, so that the build stays a small part of the work
    const int64_t tiles = (M + PIB_TILE - 1) / PIB_TILE;
    int64_t split = 1;
    if (B < 2 * 148 * 2) split = (2 * 148 * 2 + B - 1) / B;
    int64_t tiles_per_cta = (tiles + split - 1) / split;
    if (tiles_per_cta < 4) tiles_per_cta = tiles < 4 ? tiles : 4;
    if (tiles_per_cta > 64) tiles_per_cta = 64;
    const int64_t pts_per_cta = tiles_per_cta * PIB_TILE;
    const int64_t gx = (M + pts_per_cta - 1) / pts_per_cta;
    if (gx > 0x7fffffffLL) {
        set_error("num_points=%lld exceeds the grid limit", (long long)M);
        return LG_ERR_TOO_LARGE;
    }
    const size_t smem = PibSmem::total(T);
    int rc;
    for (int b0 = 0; b0 < B; b0 += 65535) {  // grid.y holds at most 65,535 frames: larger batches go out in slices
        const int bn = B - b0 < 65535 ? B - b0 : 65535;
        dim3 grid((unsigned)gx, (unsigned)bn);
        const float* bx = boxes + (size_t)b0 * T * 7;
        const float* pp = pts + (size_t)b0 * M * 3;
        int32_t* oo = out + (size_t)b0 * M;
        if (flags & LG_FLAG_STRICT_FP32) {
            if ((rc = set_smem(pib_grid_kernel<0>, smem))) return rc;
            pib_grid_kernel<0><<<grid, PIB_THREADS, smem, st>>>(bx, pp, oo, T, M, pts_per_cta);
        } else {
            if ((rc = set_smem(pib_grid_kernel<1>, smem))) return rc;
            pib_grid_kernel<1><<<grid, PIB_THREADS, smem, st>>>(bx, pp, oo, T, M, pts_per_cta);
        }
        if ((rc = check_launch("pib_grid_kernel"))) return rc;
    }
    return LG_OK;
}

extern "C" int lg_points_in_boxes_mask(const float* boxes, int64_t n, const float* pts, int64_t m, int32_t* out, float margin,
                                       unsigned flags, void* stream) {
    using namespace lg;
    if (n < 0 || m < 0) {
        set_error("negative size n=%lld m=%lld", (long long)n, (long long)m);
        return LG_ERR_INVALID_ARG;
    }
    if (n == 0 || m == 0) return LG_OK;
    if (!boxes || !pts || !out) {
        set_error("null pointer (boxes=%p pts=%p out=%p)", (const void*)boxes, (const void*)pts, (void*)out);
        return LG_ERR_INVALID_ARG;
    }
    const int64_t gy = (n + PIBM_BOXES - 1) / PIBM_BOXES;
    if (gy > 65535) {
        set_error("n=%lld boxes exceed the grid limit; split the call", (long long)n);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((m + PIB_THREADS - 1) / PIB_THREADS), (unsigned)gy);
    if (flags & LG_FLAG_STRICT_FP32) pib_mask_kernel<0><<<grid, PIB_THREADS, 0, st>>>(boxes, n, pts, m, out, margin);
    else pib_mask_kernel<1><<<grid, PIB_THREADS, 0, st>>>(boxes, n, pts, m, out, margin);
    return check_launch("pib_mask_kernel");
}
