// lg_nms.cu -- batched rotated / axis-aligned NMS, entirely on the device, for sm_100a.
//
// Replaces nms_kernel / nms_normal_kernel, their launchers and the host halves of nms_gpu /
// nms_normal_gpu (/root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:267-372, 401-414;
// iou3d_nms.cpp:90-186): cudaMalloc -> kernel -> blocking 2 MiB D2H -> cudaFree -> serial CPU sweep.
//
// Design:
//   prep   one thread per (problem, box): gather through `order`, build the 80-byte record.
//   mask   one CTA per UPPER-TRIANGLE 64x64 tile (the reference also computes the lower triangle,
//          which its sweep never reads): circle cull -> smem queue -> polygon path at full lane
//          occupancy -> 64 suppression words assembled in smem (atomicOr on 32-bit halves) ->
//          one 8-byte store per row.  mask[(p*nmax + row) * cbk + col_block].
//   sweep  one warp per problem, on the device: the 64 boxes of a diagonal block are resolved from
//          the diagonal words held in registers (shuffles, no memory in the serial chain); the rows of
//          the boxes that survive are then OR-reduced into the running `remv` words with independent,
//          batched loads.  keep[] / num_keep[] are written on the device, mapped through `order`.
// All P problems of a batch go through three launches in total, with no host synchronisation.
#include <cooperative_groups.h>

#include <atomic>

#include "lg_common.cuh"
#include "lg_strip.cuh"

namespace cg = cooperative_groups;

namespace lg {

constexpr int NMS_THREADS = 256;
constexpr int NMS_TILE = 64;
constexpr size_t NMS_STATS_BYTES = 256;  // three u64 counters (pairs cull-tested, pairs through the polygon path, pairs with non-zero overlap), padded

__device__ __forceinline__ int problem_count(const int32_t* __restrict__ counts, int p, int nmax) {
    int n = counts ? counts[p] : nmax;
    return max(0, min(n, nmax));
}

template <int FL>
__global__ void __launch_bounds__(256) nms_prep_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ order,
                                                       const int32_t* __restrict__ counts, int nmax,
                                                       float4* __restrict__ rec) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= problem_count(counts, p, nmax)) return;
    const int64_t base = (int64_t)p * nmax;
    int64_t src = i;
    if (order) {
        src = order[base + i];
        if (src < 0 || src >= nmax) src = i;  // defensive: never read out of the problem's rows
    }
    make_record<FL>(boxes + (base + src) * 7, rec + (base + i) * REC_F4);
}

// linear index over the upper triangle (row-major, cb >= rb) -> (rb, cb)
__device__ __forceinline__ void tri_decode(int t, int nb, int& rb, int& cb) {
    // offset(rb) = rb*nb - rb*(rb-1)/2
    const float fn = (float)nb + 0.5f;
    int r = (int)(fn - sqrtf(fmaxf(fn * fn - 2.0f * (float)t, 0.f)));
    r = max(0, min(r, nb - 1));
    while (r > 0 && r * nb - r * (r - 1) / 2 > t) r--;
    while (r + 1 < nb && (r + 1) * nb - (r + 1) * r / 2 <= t) r++;
    rb = r;
    cb = r + (t - (r * nb - r * (r - 1) / 2));
}

struct NmsSmem {
    static constexpr size_t rec_bytes = (size_t)2 * NMS_TILE * REC_F4 * sizeof(float4);
    static constexpr size_t mask_bytes = (size_t)NMS_TILE * sizeof(unsigned long long);
    static constexpr size_t total = rec_bytes + ST_SLAB_BYTES + (size_t)(ST_QCAP + ST_RARECAP) * sizeof(uint16_t) + mask_bytes;
};

template <int FL>
__global__ void __launch_bounds__(NMS_THREADS, 3)
    nms_mask_kernel(const float4* __restrict__ rec, const int32_t* __restrict__ counts, const int nmax, const int cbk,
                    const float thresh, unsigned long long* __restrict__ mask) {
    constexpr int NT = NMS_THREADS, T = NMS_TILE;
    static_assert(NT == ST_THREADS && T * T <= ST_QCAP, "queue sizing");
    extern __shared__ float4 smem4[];
    float4* sA = smem4;
    float4* sB = sA + T * REC_F4;
    float2* slab = reinterpret_cast<float2*>(sB + T * REC_F4);
    uint16_t* queue = reinterpret_cast<uint16_t*>(slab + SLAB_ROWS_SMEM_B * NT);
    uint16_t* rareq = queue + ST_QCAP;
    unsigned int* smask = reinterpret_cast<unsigned int*>(rareq + ST_RARECAP);  // 64 x (lo, hi)
    __shared__ int qcount, rcount;

    const int p = blockIdx.y;
    const int n = problem_count(counts, p, nmax);
    const int nb = (n + T - 1) / T;
    int rb, cb;
    tri_decode(blockIdx.x, cbk, rb, cb);
    if (rb >= nb || cb >= nb) return;  // tile outside this problem's boxes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int na = min(T, n - rb * T), ncol = min(T, n - cb * T);
    const float4* gA = rec + ((int64_t)p * nmax + (int64_t)rb * T) * REC_F4;
    const float4* gB = rec + ((int64_t)p * nmax + (int64_t)cb * T) * REC_F4;
    for (int e = tid; e < na * REC_F4; e += NT) sA[e] = __ldg(gA + e);
    for (int e = tid; e < ncol * REC_F4; e += NT) sB[e] = __ldg(gB + e);
    if (tid < 2 * T) smask[tid] = 0u;
    if (tid == 0) {
        qcount = 0;
        rcount = 0;
    }
    __syncthreads();

    {
        const int col = (warp & 1) * 32 + lane, rsub = warp >> 1;
        const bool diag = (rb == cb);
        const bool cvalid = col < ncol;
        const float4 bc = cvalid ? sB[col * REC_F4 + REC_CULL] : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned mk[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int r = rsub + 4 * k;
            bool surv = false;
            if (cvalid && r < na && (!diag || col > r)) surv = cull_survives(sA[r * REC_F4 + REC_CULL], bc);
            mk[k] = __ballot_sync(0xffffffffu, surv);
        }
        push_survivors<8, 16>(mk, lane, rsub, 4, col, &qcount, queue);
    }
    __syncthreads();

    // row = the higher-scoring box: iou_bev(row, col), kernel.cu:304
    auto emit = [&](int r, int c, float ov, const float4* A, const float4* B) {
        const float iou = iou_from_overlap(ov, A[REC_CULL].w, B[REC_CULL].w);
        if (iou > thresh) atomicOr(&smask[2 * r + (c >> 5)], 1u << (c & 31));
    };
    drain_main<FL, 8>(sA, sB, slab, queue, qcount, rareq, &rcount, emit);
    __syncthreads();
    drain_rare<FL, 8>(sA, sB, slab, rareq, &rcount, emit);
    __syncthreads();

    if (tid < na) {
        const unsigned long long w = (unsigned long long)smask[2 * tid] | ((unsigned long long)smask[2 * tid + 1] << 32);
        mask[((int64_t)p * nmax + (int64_t)rb * T + tid) * cbk + cb] = w;
    }
}

// axis-aligned variant: the pair test is ~25 instructions, so one thread per row, like the reference,
// but upper-triangle tiles only and batched.
template <int FL>
__global__ void __launch_bounds__(NMS_TILE)
    nms_normal_mask_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ order,
                           const int32_t* __restrict__ counts, const int nmax, const int cbk, const float thresh,
                           unsigned long long* __restrict__ mask) {
    constexpr int T = NMS_TILE;
    __shared__ float4 sB[T];
    const int p = blockIdx.y;
    const int n = problem_count(counts, p, nmax);
    const int nb = (n + T - 1) / T;
    int rb, cb;
    tri_decode(blockIdx.x, cbk, rb, cb);
    if (rb >= nb || cb >= nb) return;
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)p * nmax;
    auto load = [&](int i) {
        int64_t src = i;
        if (order) {
            src = order[base + i];
            if (src < 0 || src >= nmax) src = i;
        }
        const float* b = boxes + (base + src) * 7;
        return make_float4(b[0], b[1], b[3], b[4]);
    };
    const int ncol = min(T, n - cb * T), na = min(T, n - rb * T);
    if (tid < ncol) sB[tid] = load(cb * T + tid);
    __syncthreads();
    if (tid < na) {
        const float4 a = load(rb * T + tid);
        unsigned long long w = 0;
        const int start = (rb == cb) ? tid + 1 : 0;
        for (int c = start; c < ncol; c++) {
            const float4 b = sB[c];
            if (iou_normal<FL>(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w) > thresh) w |= 1ull << c;
        }
        mask[(base + (int64_t)rb * T + tid) * cbk + cb] = w;
    }
}

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}

// one warp per problem (iou3d_nms.cpp:116-132 restated for the device)
__global__ void __launch_bounds__(32)
    nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int64_t* __restrict__ order,
                     const int32_t* __restrict__ counts, const int nmax, const int cbk, int64_t* __restrict__ keep, const int keep_ld,
                     const int max_keep, int32_t* __restrict__ num_keep) {
    extern __shared__ unsigned long long remv[];
    const int p = blockIdx.x, lane = threadIdx.x;
    const int n = problem_count(counts, p, nmax);
    const int nb = (n + 63) / 64;
    const int64_t base = (int64_t)p * nmax;
    const unsigned long long* M = mask + base * cbk;
    int64_t* const krow = keep + (int64_t)p * keep_ld;
    for (int w = lane; w < nb; w += 32) remv[w] = 0ull;
    __syncwarp();
    int nk = 0;
    for (int b = 0; b < nb && nk < max_keep; b++) {  // NMS_POST_MAXSIZE reached: the rest cannot be kept
        const int r0 = b * 64 + lane, r1 = r0 + 32;
        const unsigned long long d0 = (r0 < n) ? M[(int64_t)r0 * cbk + b] : 0ull;
        const unsigned long long d1 = (r1 < n) ? M[(int64_t)r1 * cbk + b] : 0ull;
        unsigned long long cur = remv[b];
        const int valid = min(64, n - b * 64);
        if (valid < 64) cur |= ~0ull << valid;  // rows past the end are never kept
        unsigned long long kept = 0ull;
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            const unsigned long long di = shfl64(d0, i);
            if (!((cur >> i) & 1ull)) {
                kept |= 1ull << i;
                cur |= di;
            }
        }
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            const unsigned long long di = shfl64(d1, i);
            if (!((cur >> (i + 32)) & 1ull)) {
                kept |= 1ull << (i + 32);
                cur |= di;
            }
        }
        // emit kept indices in order
        {
            const unsigned long long below0 = kept & ((1ull << lane) - 1ull);
            const unsigned long long below1 = kept & ((1ull << (lane + 32)) - 1ull);
            const int q0 = nk + __popcll(below0), q1 = nk + __popcll(below1);
            if (((kept >> lane) & 1ull) && q0 < max_keep) krow[q0] = order ? order[base + r0] : (int64_t)r0;
            if (((kept >> (lane + 32)) & 1ull) && q1 < max_keep) krow[q1] = order ? order[base + r1] : (int64_t)r1;
        }
        nk = min(nk + __popcll(kept), max_keep);
        // fold the kept rows into remv for the later column blocks (independent loads, 4 in flight)
        for (int w = b + 1 + lane; w < nb; w += 32) {
            unsigned long long acc = remv[w];
            unsigned long long kk = kept;
            while (kk) {
                unsigned long long v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    v[u] = 0ull;
                    if (kk) {
                        const int i = __ffsll((long long)kk) - 1;
                        kk &= kk - 1ull;
                        v[u] = M[(int64_t)(b * 64 + i) * cbk + w];
                    }
                }
                acc |= v[0] | v[1] | v[2] | v[3];
            }
            remv[w] = acc;
        }
        __syncwarp();
    }
    if (lane == 0) num_keep[p] = nk;
    for (int i = nk + lane; i < max_keep; i += 32) krow[i] = -1;
}

// ---------------------------------------------------------------------------------------------------
// Lazy rotated NMS: one CTA (or one thread-block cluster) per problem, no N x N/64 mask, no separate record kernel.
//
// The reference's sweep (iou3d_nms.cpp:121-132) reads row i of the mask only when box i is KEPT, so only
// kept rows have to exist.  A problem is worked off in passes; every pass decides a set of CANDIDATES taken from the
// first alive (undecided) boxes in score order, and evaluates their rows against every later alive box at once:
//   select -> the alive boxes at and after the cursor are listed densely; the first 4 G of them are the WINDOW.
//             independent mode (the default): a window box is a candidate iff the exact-zero circle test separates it from
//               EVERY earlier window box -- whatever those turn out to be, none of them can suppress it, and everything before
//               the window is decided, so a candidate is KEPT, unconditionally; up to G candidates per pass, no speculation.
//               The others wait: most are suppressed by a candidate in this very pass, the rest head the next window.
//             speculative mode (high thresholds, where overlapping boxes mostly survive): the first G alive boxes are the
//               candidates, all assumed kept; the resolve below sorts it out.  A problem starts in this mode when
//               thresh > 0.3 and leaves it for good as soon as a pass loses more than a quarter of its candidates.
//   cull   -> (candidate, box) codes of the pairs the circle test cannot prove disjoint, in a smem queue;
//   drain  -> the polygon path on full warps; iou(candidate, box) > thresh marks the box: in the pass's kill bitmap
//             (independent mode) or in the candidate's suppression row (speculative mode);
//   resolve-> speculative mode only: candidates are walked in score order, one that an earlier KEPT candidate of the same
//             pass suppresses is dropped together with its row; the rows of the others are OR-ed into the kill words;
//   kill   -> the marked boxes leave the alive bitmap.
// Every IoU that decides anything is iou_bev(kept box, later box) exactly as in the mask formulation, so the
// keep list is identical; the work drops from N^2/2 pairs to about #kept x #alive.
//
// A problem is a chain of dependent passes, each of them a handful of latency-bound phases (one polygon round is a
// ~1500-instruction dependency chain per lane), so the kernel is as fast as it has few passes and few rounds per pass.
constexpr int LZ_THREADS = 512;  // one problem is latency-bound: 16 warps hide the polygon path's dependency chains
constexpr int LZ_G = 32;         // speculative candidates per pass (one lane each in the resolve, one suppression row each)
constexpr int LZ_GI = 32;        // independent candidates per pass (they need no rows and no resolve).  64 (window 256) was measured: the
                                 // window of the best-scored alive boxes holds ~35 mutually separated ones, so the number of passes barely
                                 // moves (4.3 -> 4.1 on nms_cfg2) while the window's conflict tests quadruple (kernel 0.153 -> 0.170 ms)
constexpr int LZ_GH = 16;        // candidates per warp in a sweep: the warps form ceil(candidates / 16) groups that share every 32 columns
constexpr int LZ_WIN = 4 * LZ_GI; // window: the first alive boxes among which the candidates are chosen
constexpr int LZ_CACHE = 4096;   // cull quads cached in smem; boxes beyond read theirs from the records (L2)
constexpr int LZ_MAX_CLUSTER = 8;     // portable cluster size limit
constexpr size_t LZ_SMEM_LIMIT = 227 * 1024 - 2048;  // opt-in shared memory per CTA on sm_100a (227 KB), minus the kernel's static part (1.6 KB)

struct LazyLayout {
    int words, sstride;  // alive words; pitch of a suppression row (odd: the resolve reads a column of the rows conflict-free)
    int qcap, rarecap;   // u32 codes: candidate << 16 | box.  qcap = one sweep step of the CTA at 100 % survivors + all candidate pairs
    int gcap;            // candidates per pass: LZ_G, fewer when the suppression rows of a very large problem would not fit
    size_t off_cand, off_cull, off_slab, off_queue, off_rare, off_alive, off_kill, off_kept, off_prefix, off_win, off_conf, off_dlist, off_sup, total;
    __host__ __device__ explicit LazyLayout(int nmax, int nt = LZ_THREADS, int gcap_ = LZ_G) {
        words = (nmax + 31) / 32;
        sstride = words | 1;
        qcap = 16 * nt + 512;
        rarecap = qcap + 1024;
        gcap = gcap_;
        size_t o = (size_t)LZ_GI * REC_F4 * sizeof(float4);  // candidate records first
        off_cand = o;
        o += (size_t)LZ_GI * sizeof(float4);
        off_cull = o;
        o += (size_t)(nmax < LZ_CACHE ? nmax : LZ_CACHE) * sizeof(float4);
        off_slab = o;
        o += (size_t)SLAB_ROWS * nt * sizeof(float2);
        off_queue = o;
        o += (size_t)qcap * sizeof(uint32_t);
        off_rare = o;
        o += (size_t)rarecap * sizeof(uint32_t);
        off_alive = o;
        o += (size_t)words * sizeof(uint32_t);
        off_kill = o;
        o += (size_t)words * sizeof(uint32_t);
        off_kept = o;
        o += (size_t)words * sizeof(uint32_t);
        off_prefix = o;
        o += (size_t)words * sizeof(int);
        off_win = o;
        o += (size_t)LZ_WIN * sizeof(int);
        off_conf = o;
        o += (size_t)LZ_WIN * sizeof(int);
        off_dlist = o;
        o += ((size_t)nmax * sizeof(uint16_t) + 15) / 16 * 16;
        off_sup = o;
        o += (size_t)gcap * sstride * sizeof(uint32_t);
        total = o;
    }
};

static inline int lazy_gcap(int nmax, int nt) {  // largest of 32 / 16 / 8 candidates per pass whose layout fits, 0 if none does
    for (int g = LZ_G; g >= 8; g >>= 1)
        if (LazyLayout(nmax, nt, g).total <= LZ_SMEM_LIMIT) return g;
    return 0;
}

// Fused gather (lg_nms_rotated_gather): the keep list of problem p goes, packed as (count, kept indices..., -1...), into row
// row0 + p of a (rows, 1 + max_keep) int64 buffer on EVERY rank of the job -- plain stores through the NVLink peer mappings
// (CUDA peer access / symmetric memory), straight from the kernel's epilogue, instead of a keep tensor + a collective.
struct PeerRows {
    int64_t* buf[LG_MAX_PEERS];
    int n;         // 0: not a gather call
    int64_t row0;  // this rank's first row in every buffer
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// NT = 512 threads for one latency-bound problem per SM; NT = 256 (two CTAs per SM) when the batch has several small
// problems per SM, so that one problem's serial phases (selection, resolve, barriers) overlap another's rounds.
// rec: workspace, written by the kernel's own prologue (the 112-byte records of lg_geom.cuh, in score order) and read
// back through L2 -- never through the read-only path, which is not coherent with writes of the same launch.
template <int FL, bool CL, int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 1 : 2)
    nms_lazy_kernel(const float* __restrict__ boxes, float4* rec, const int64_t* __restrict__ order, const int32_t* __restrict__ counts,
                    const int nmax, const float thresh, int64_t* __restrict__ keep, const int keep_ld, const int max_keep,
                    int32_t* __restrict__ num_keep, unsigned long long* __restrict__ stats, const int gcap, const PeerRows peers) {
    constexpr int G = LZ_G, GH = LZ_GH, NW = NT / 32;
    extern __shared__ float4 smem4[];
    const LazyLayout L(nmax, NT, gcap);
    char* sm = reinterpret_cast<char*>(smem4);
    float4* sA = smem4;
    float4* scand = reinterpret_cast<float4*>(sm + L.off_cand);
    float4* scull = reinterpret_cast<float4*>(sm + L.off_cull);
    float2* slab = reinterpret_cast<float2*>(sm + L.off_slab);
    uint32_t* queue = reinterpret_cast<uint32_t*>(sm + L.off_queue);
    uint32_t* rareq = reinterpret_cast<uint32_t*>(sm + L.off_rare);
    uint32_t* alive = reinterpret_cast<uint32_t*>(sm + L.off_alive);
    uint32_t* killw = reinterpret_cast<uint32_t*>(sm + L.off_kill);
    uint32_t* keptw = reinterpret_cast<uint32_t*>(sm + L.off_kept);  // kept boxes: independent candidates are kept out of score order
    int* wprefix = reinterpret_cast<int*>(sm + L.off_prefix);
    int* window = reinterpret_cast<int*>(sm + L.off_win);
    int* wconf = reinterpret_cast<int*>(sm + L.off_conf);
    uint16_t* dlist = reinterpret_cast<uint16_t*>(sm + L.off_dlist);
    uint32_t* sup = reinterpret_cast<uint32_t*>(sm + L.off_sup);
    const int SW = L.sstride, QCAP = L.qcap, RARECAP = L.rarecap;
    __shared__ int qcount, rcount, group[LZ_GI], ng_s, keptmask_s, nk_s, qvalid_s, sfail_s, nwin_s, nalive_s, cursor_s, spec_s;
    // the candidates' cull quads once more, two candidates side by side: (x0, x1, y0, y1) and (r0, r1) -- the operand pairs of the
    // packed FP32 instructions of the sweeps
    __shared__ int wtot[NT / 32], full_s;  // the alive scan: totals of the warps' 32-word slices; NMS_POST_MAXSIZE reached
    __shared__ float4 sc2xy[LZ_GI / 2];
    __shared__ float2 sc2r[LZ_GI / 2];
    __shared__ unsigned long long st_heavy;

    // A problem may be spread over a thread-block cluster of C CTAs (C SMs): they keep identical copies of the alive bitmap
    // and of the pass state, split the COLUMNS of every pass (interleaved sweep steps), evaluate the few
    // candidate-vs-candidate pairs redundantly, and exchange the kill words through distributed shared memory.
    // (CL = false compiles all of that out: one CTA per problem.)
    cg::cluster_group cluster = cg::this_cluster();
    const int C = CL ? (int)cluster.num_blocks() : 1, crank = CL ? (int)cluster.block_rank() : 0;
    const bool split = CL && C > 1;  // the problem's columns are split over several CTAs
    const int p = blockIdx.x / C;
    const int n = problem_count(counts, p, nmax);
    const int W = (n + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)p * nmax;
    float4* grec = rec + base * REC_F4;

#ifdef LG_LZ_TIMING  // developer build: cycles per phase of thread 0, accumulated into stats[8 + phase] (tools/lz_timing.py)
    long long tmark = clock64();
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int npass = 0;
#define LZ_MARK(ph)                      \
    if (tid == 0) {                      \
        const long long now_ = clock64(); \
        tacc[ph] += now_ - tmark;        \
        tmark = now_;                    \
    }
#else
#define LZ_MARK(ph)
#endif
    // ---- prologue: the problem's records, in score order (every CTA of a cluster builds an interleaved share)
    // The chunks of NT boxes are dealt to the CTAs in turn, 32 boxes of a chunk to every warp; the warps work on their own, no CTA
    // barrier: (1) the warp gathers its 32 boxes (7 floats each, through `order`) with flat loads -- lane l takes floats l, l + 32,
    // ... of the 224, so a warp instruction touches ~6 cache lines instead of the 32 of one-box-per-lane loads; the next chunk's
    // boxes are requested before this chunk's arithmetic starts; (2) one record per lane, built in shared memory; (3) the warp's
    // 32 records (3584 contiguous bytes) leave as one stream.  (One box and one record per lane straight from / to global memory
    // cost 64 cache-line visits per warp instruction pair: 18 us of a 150 us problem.)  Staging: the slab + the queues, idle until
    // the first pass.
    {
        float* wraw = reinterpret_cast<float*>(sm + L.off_slab) + warp * (32 * 7);
        float4* wrecs = reinterpret_cast<float4*>(sm + L.off_slab + (size_t)NT * 7 * sizeof(float)) + warp * (32 * REC_F4);
        const int S = C * NT;
        auto load_src = [&](const int i) -> int64_t {
            int64_t src = i < n ? i : 0;
            if (order && i < n) {
                const int64_t sv = __ldg(order + base + i);
                if (sv >= 0 && sv < nmax) src = sv;  // defensive: never read out of the problem's rows
            }
            return src;
        };
        float v[7];
        auto gather = [&](const int64_t src) {
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const int e = lane + 32 * k;
                const int64_t sb = __shfl_sync(0xffffffffu, src, e / 7);
                v[k] = __ldg(boxes + (base + sb) * 7 + (e % 7));
            }
        };
        const int c00 = crank * NT;
        int64_t src_next = load_src(c00 + S + tid);
        if (c00 < n) gather(load_src(c00 + tid));
        for (int c0 = c00; c0 < n; c0 += S) {
            const int i = c0 + tid;
#pragma unroll
            for (int k = 0; k < 7; k++) wraw[lane + 32 * k] = v[k];
            __syncwarp();
            if (c0 + S < n) gather(src_next);            // in flight during this chunk's arithmetic
            src_next = load_src(c0 + 2 * S + tid);
            if (i < n) {
                const float* b = wraw + lane * 7;
                float4* r = wrecs + lane * REC_F4;
                make_record_v<FL>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], r);
                if (i < LZ_CACHE) scull[i] = r[REC_CULL];
            }
            __syncwarp();
            const int cnt = min(32, n - (c0 + warp * 32));  // records of this warp in this chunk (<= 0: none)
            float4* gdst = grec + (int64_t)(c0 + warp * 32) * REC_F4;
            for (int e = lane; e < cnt * REC_F4; e += 32) gdst[e] = wrecs[e];
            __syncwarp();
        }
    }
    for (int w = tid; w < W; w += NT) {
        alive[w] = (w == W - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xFFFFFFFFu;
        killw[w] = 0u;
        keptw[w] = 0u;
    }
    for (int w = tid; w < gcap * SW; w += NT) sup[w] = 0u;
    if (tid == 0) {
        qcount = 0;
        rcount = 0;
        nk_s = 0;
        qvalid_s = QCAP;
        sfail_s = 0x7fffffff;
        cursor_s = 0;
        spec_s = thresh > 0.3f ? 1 : 0;
        st_heavy = 0ull;
    }
    // records visible to the whole problem (cluster barrier: release / acquire at cluster scope); it also tells every CTA that
    // its peers have started and initialised their bitmaps, which must hold before any of them is touched remotely
    LZ_MARK(9)  // prologue: records
    if (split) cluster.sync();
    else __syncthreads();
    LZ_MARK(10)  // prologue: barrier
    if (split) {  // the cull quads of the chunks the peers built (this CTA's own went straight into scull)
        for (int j = tid; j < min(n, LZ_CACHE); j += NT)
            if ((j / NT) % C != crank) scull[j] = __ldcg(grec + (int64_t)j * REC_F4 + REC_CULL);
    }
    unsigned my_tested = 0u, my_nonzero = 0u;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    auto quad = [&](const int j) -> float4 { return j < LZ_CACHE ? scull[j] : __ldcg(grec + (int64_t)j * REC_F4 + REC_CULL); };

    // optimistic push: the sweeps of a pass run without a barrier and reserve queue space as they go; a reservation that does
    // not fit writes nothing and records where the queue stopped being complete (every later reservation fails as well, so the
    // entries below qvalid_s are exactly those of the successful ones) and the earliest step that lost pairs.
    auto push = [&](const unsigned (&mk)[GH], const int rbase, const int col, const int step) {
        unsigned any = 0u;
#pragma unroll
        for (int k = 0; k < GH; k++) any |= mk[k];
        if (any == 0u) return;
        int total = 0;
#pragma unroll
        for (int k = 0; k < GH; k++) total += __popc(mk[k]);
        int qb = 0;
        if (lane == 0) {
            qb = atomicAdd(&qcount, total);
            if (qb + total > QCAP) {
                atomicMin(&qvalid_s, qb);
                atomicMin(&sfail_s, step);
            }
        }
        qb = __shfl_sync(FULL, qb, 0);
        if (qb + total > QCAP) return;
#pragma unroll
        for (int k = 0; k < GH; k++) {
            if (mk[k]) {
                if ((mk[k] >> lane) & 1u) queue[qb + __popc(mk[k] & lt)] = (uint32_t)(((rbase + k) << 16) | col);
                qb += __popc(mk[k]);
            }
        }
    };

    while (true) {
        __syncthreads();
        const int cursor = cursor_s;  // every box below it is decided
        const bool spec = spec_s != 0;
        const int gmax = spec ? gcap : min(LZ_GI, 2 * gcap);  // candidates of this pass at most
        const int wincap = spec ? gcap : 4 * gmax;
        // ---- (all warps) alive boxes at and after the cursor: per-word offsets of the dense list, and the window.  Warp k takes the
        // k-th 32 words of a round of NW x 32 (one round up to 16,384 boxes); the warps' totals meet in shared memory.  (One warp
        // walking the bitmap 32 words at a time while fifteen waited cost 2 us of every 21-us pass.)
        if (warp == 0) {
            // NMS_POST_MAXSIZE (model_nms_utils.py:20): the first max_keep kept boxes in score order are final once that many
            // are kept BELOW the cursor (independent candidates are kept out of order; everything below the cursor is decided)
            bool full = max_keep <= 0;
            if (!full && nk_s >= max_keep) {
                int kb = 0;
                for (int w = lane; w <= (cursor >> 5) && w < W; w += 32) {
                    unsigned word = keptw[w];
                    if (w == (cursor >> 5)) word &= ~(FULL << (cursor & 31));
                    kb += __popc(word);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) kb += __shfl_xor_sync(FULL, kb, d);
                full = kb >= max_keep;
            }
            if (lane == 0) full_s = full ? 1 : 0;
        }
        int total = 0;  // the same in every thread
        for (int g0 = cursor >> 5; g0 < W; g0 += NW * 32) {
            const int w = g0 + warp * 32 + lane;
            unsigned word = w < W ? alive[w] : 0u;
            if (w == (cursor >> 5)) word &= FULL << (cursor & 31);
            const int c = __popc(word);
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) wtot[warp] = incl;
            __syncthreads();
            int ws = lane < NW ? wtot[lane] : 0;  // inclusive prefix of the warps' totals
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, ws, d);
                if (lane >= d) ws += v;
            }
            const int before = warp > 0 ? __shfl_sync(FULL, ws, warp - 1) : 0, round_total = __shfl_sync(FULL, ws, NW - 1);
            int slot = total + before + incl - c;
            if (w < W) wprefix[w] = slot;
            while (word && slot < wincap) {  // the window: every lane lists the boxes of its own word (few, once the bitmap thins out)
                const int b = __ffs(word) - 1;
                word &= word - 1;
                window[slot++] = w * 32 + b;
            }
            total += round_total;
            if (g0 + NW * 32 < W) __syncthreads();  // wtot is written again in the next round
        }
        if (tid == 0) {
            if (full_s) total = 0;  // (warp 0 wrote full_s before the round's barrier; a problem without a round has no boxes left)
            nalive_s = total;
            nwin_s = min(total, wincap);
            qcount = 0;  // the queue of the previous pass is drained
            qvalid_s = QCAP;
            sfail_s = 0x7fffffff;
        }
        __syncthreads();
        LZ_MARK(0)  // alive scan
        const int nal = nalive_s, nwin = nwin_s;
        if (nal == 0) break;  // every CTA of a cluster derives the same pass from identical bitmaps: all leave together
        // ---- (all) the dense list of alive boxes, four threads per word; and, in independent mode, which window boxes the
        // circle test cannot separate from an earlier window box
        for (int wb = cursor >> 5; wb < W; wb += NT / 4) {
            const int w = wb + (tid >> 2), part = tid & 3;
            if (w < W) {
                unsigned word = alive[w];
                if (w == (cursor >> 5)) word &= FULL << (cursor & 31);
                unsigned bits = (word >> (8 * part)) & 0xffu;
                int off = wprefix[w] + __popc(word & ((1u << (8 * part)) - 1u));
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    dlist[off++] = (uint16_t)(w * 32 + 8 * part + b);
                }
            }
        }
        if (!spec) {
            const int chunk = (wincap + 3) >> 2;
            for (int b0 = 0; b0 < nwin; b0 += NT / 4) {
                const int b = b0 + (tid >> 2), part = tid & 3;
                bool cf = false;
                if (b < nwin) {
                    const float4 qb = quad(window[b]);
                    const int a1 = min(part * chunk + chunk, b);
                    for (int a = part * chunk; a < a1; a++) cf |= cull_survives(quad(window[a]), qb);
                    my_tested += (unsigned)max(a1 - part * chunk, 0);
                }
                const unsigned m = __ballot_sync(FULL, cf);
                if (b < nwin && part == 0) wconf[b] = (int)((m >> (lane & ~3)) & 0xfu);
            }
        }
        __syncthreads();
        LZ_MARK(1)  // dense list + window conflicts
        // ---- (warp 0) the candidates
        if (warp == 0) {
            int cnt = 0, firstskip = -1;
            for (int c0 = 0; c0 < nwin; c0 += 32) {
                const int b = c0 + lane;
                const bool inw = b < nwin;
                const bool f = inw && (spec || wconf[b] == 0);
                const unsigned m = __ballot_sync(FULL, f);
                const int slot = cnt + __popc(m & lt);
                const bool sel = f && slot < gmax;
                if (sel) group[slot] = window[b];
                const unsigned un = __ballot_sync(FULL, inw && !sel);
                if (firstskip < 0 && un) firstskip = c0 + __ffs(un) - 1;
                cnt += __popc(m);
            }
            if (lane == 0) {
                ng_s = min(cnt, gmax);  // >= 1: the first window box has no earlier one
                // next cursor: the first window box that is not a candidate, else the first alive box after the window
                cursor_s = firstskip >= 0 ? window[firstskip] : (nwin < nal ? (int)dlist[nwin] : n);
            }
        }
        __syncthreads();
        const int ng = ng_s;
        for (int e = tid; e < ng * REC_F4; e += NT) sA[e] = __ldcg(grec + (int64_t)group[e / REC_F4] * REC_F4 + (e % REC_F4));
        if (tid < ng) {
            const int j = group[tid];
            const float4 qj = quad(j);
            scand[tid] = qj;
            reinterpret_cast<float*>(&sc2xy[tid >> 1])[tid & 1] = qj.x;
            reinterpret_cast<float*>(&sc2xy[tid >> 1])[2 + (tid & 1)] = qj.y;
            reinterpret_cast<float*>(&sc2r[tid >> 1])[tid & 1] = qj.z;
            // the candidates are decided in this pass: they leave the alive bitmap now (every CTA of a cluster clears its own
            // copy), so the column sweeps below never see them -- candidate-vs-candidate pairs are queued separately
            atomicAnd(&alive[j >> 5], ~(1u << (j & 31)));
        }
        __syncthreads();
        LZ_MARK(2)  // candidates + their records
        const int g0 = group[0];
        auto emit = [&](int g, int j, float ov, const float4* A, const float4* B) {
            const float iou = iou_from_overlap(ov, A[REC_CULL].w, B[REC_CULL].w);  // row = the higher-scoring box (kernel.cu:304)
            my_nonzero += ov > 0.f ? 1u : 0u;
            if (iou > thresh) atomicOr(spec ? &sup[g * SW + (j >> 5)] : &killw[j >> 5], 1u << (j & 31));
        };
        // ---- rows of the candidates against every later alive box.  The warps form ceil(ng / GH) groups; warp group `half` tests
        // its GH candidates against 32 entries of the dense list per step (a step of the CTA: STEP columns x all candidates, at
        // most NW x 512 pairs -- the queue's capacity); the steps of a pass are interleaved over the CTAs of a cluster.
        const int ngr = (ng + GH - 1) / GH, WG = NW / ngr, STEP = WG * 32;
        const int half = warp / WG, wcol = warp % WG;
        const int kmax = half < ngr ? min(GH, ng - GH * half) : 0;  // candidates of this warp group in this pass (0: an idle warp)
        const int ghi = kmax > 0 ? group[GH * half + kmax - 1] : 0;  // the last of them
        const int nsteps = (nal + C * STEP - 1) / (C * STEP);  // per CTA
        int s0 = 0;          // first step not yet swept
        bool safe = false;   // after an overflow: one step at a time, which an empty queue always holds
        while (true) {
            const int s1 = safe ? min(s0 + 1, nsteps) : nsteps;
            // candidate g against the later candidates (part of step 0; speculative mode only -- independent candidates are
            // separated by the circle test): EVERY CTA of a cluster evaluates these (at most G (G - 1) / 2) pairs itself, so
            // each has the complete candidate-vs-candidate bits for the resolve without an exchange through DSMEM
            if (spec && s0 == 0 && warp < G / GH && warp * GH < ng) {
                const bool a = lane < ng;
                const int j = a ? group[lane] : 0;
                const float4 cj = a ? scand[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                unsigned mk[GH];
#pragma unroll
                for (int k = 0; k < GH; k++) {
                    const int g = warp * GH + k;
                    mk[k] = __ballot_sync(FULL, a && lane > g && cull_survives(scand[g < ng ? g : 0], cj));  // group[] ascends: lane > g <=> a later box
                }
                push(mk, warp * GH, j, 0);
                if (lane == 0 && warp == 0 && crank == 0 && !safe) my_tested += (unsigned)(ng * (ng - 1) / 2);
            }
            if (kmax > 0) {
                for (int s = s0; s < s1; s++) {
                    const int pos = (s * C + crank) * STEP + wcol * 32 + lane;
                    const int j = pos < nal ? (int)dlist[pos] : 0;
                    const bool a = pos < nal && ((alive[j >> 5] >> (j & 31)) & 1u);  // the candidates have left the bitmap
                    if (!__any_sync(FULL, a)) continue;
                    const float4 cj = a ? quad(j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    unsigned mk[GH];
                    if (__all_sync(FULL, !a || j > ghi)) {  // the usual case: every column lies after all of this group's candidates
                        // two candidates per packed instruction (FADD2 / FMUL2 / FFMA2); the circle test is conservative, so how its
                        // products are contracted does not matter
                        const f32x2 bx = pack2(cj.x, cj.x), by = pack2(cj.y, cj.y), br = pack2(cj.z, cj.z);
#pragma unroll
                        for (int k = 0; k < GH; k += 2) {
                            mk[k] = 0u;
                            mk[k + 1] = 0u;
                            if (k < kmax) {  // warp-uniform guard
                                const ulonglong2 axy = *reinterpret_cast<const ulonglong2*>(&sc2xy[(GH * half + k) >> 1]);
                                const f32x2 ar = *reinterpret_cast<const f32x2*>(&sc2r[(GH * half + k) >> 1]);
                                const f32x2 dx = sub2(axy.x, bx), dy = sub2(axy.y, by), rr = add2(ar, br);
                                const f32x2 d2 = fma2(dx, dx, mul2(dy, dy)), r2 = mul2(rr, rr);
                                float d0, d1, r0, r1;
                                unpack2(d2, d0, d1);
                                unpack2(r2, r0, r1);
                                mk[k] = __ballot_sync(FULL, a && !(d0 > r0));  // NaN => keep: the polygon path decides
                                if (k + 1 < kmax) mk[k + 1] = __ballot_sync(FULL, a && !(d1 > r1));
                            }
                        }
                        my_tested += a ? (unsigned)kmax : 0u;
                    } else {  // window boxes that are not candidates sit between the candidates: only the earlier candidates test them
#pragma unroll
                        for (int k = 0; k < GH; k++) {
                            mk[k] = 0u;
                            if (k < kmax) {
                                const bool t = a && j > group[GH * half + k];
                                my_tested += t ? 1u : 0u;
                                mk[k] = __ballot_sync(FULL, t && cull_survives(scand[GH * half + k], cj));
                            }
                        }
                    }
                    push(mk, GH * half, j, s);
                }
            }
            __syncthreads();
            const int qn = min(qcount, qvalid_s), rn = rcount, sfail = sfail_s;
            LZ_MARK(3)  // cull sweeps
            if (rn + qn > RARECAP) {
                drain_rare<FL, 16, NT>(sA, grec, slab, rareq, &rcount, emit);
                __syncthreads();
            }
            drain_main<FL, 16, NT, true>(sA, grec, slab, queue, qn, rareq, &rcount, emit);  // column records in global memory: staged
            if (tid == 0) st_heavy += (unsigned long long)qn;
            LZ_MARK(4)  // polygon rounds (thread 0's share; the barrier that follows is charged to the next phase)
            if (sfail == 0x7fffffff && s1 >= nsteps) break;
            // overflow (adversarial inputs: nearly every pair survives the cull): go on step by step from the first step that lost
            // pairs; steps >= sfail that did fit are simply evaluated again -- marking a box twice changes nothing
            __syncthreads();  // everybody has read the counters and left the queue
            if (tid == 0) {
                qcount = 0;
                qvalid_s = QCAP;
                sfail_s = 0x7fffffff;
            }
            s0 = sfail == 0x7fffffff ? s1 : sfail;
            safe = true;
            __syncthreads();
        }
        __syncthreads();
        // cluster: this CTA has read its bitmap for the last time in this pass (the dense list and the sweeps' alive test are
        // behind it); peers wait for this arrival before their kill words of this pass may touch it (cluster_wait below)
        if (split) cluster_arrive();
        LZ_MARK(5)  // waiting for the last polygon round
        drain_rare<FL, 16, NT>(sA, grec, slab, rareq, &rcount, emit);  // the marks must be complete before the resolve / the kills
        __syncthreads();
        LZ_MARK(6)  // deferred pairs
        // ---- (warp 0, lane g = candidate g) keep list; in speculative mode first resolve the speculation in score order
        if (warp == 0) {
            const int jl = lane < ng ? group[lane] : 0;
            unsigned km = 0u;
            if (spec) {
                // colm: the earlier candidates h whose row suppresses this lane's candidate (a column of the rows: SW is odd and
                // the candidates sit in a few neighbouring words, so the loads are broadcasts or conflict-free)
                unsigned colm = 0u;
                const int wl = jl >> 5, bl = jl & 31;
                for (int h = 0; h < ng; h++) colm |= ((sup[h * SW + wl] >> bl) & 1u) << h;
                colm &= lt;
                km = 0u;
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const unsigned cg_ = __shfl_sync(FULL, colm, g);  // independent of the chain: issued ahead of it
                    if (g < ng && (cg_ & km) == 0u) km |= 1u << g;
                }
            }
            if (spec) {
                if ((km >> lane) & 1u) atomicOr(&keptw[jl >> 5], 1u << (jl & 31));
            } else {  // independent candidates (up to LZ_GI) are all kept
                for (int g = lane; g < ng; g += 32) atomicOr(&keptw[group[g] >> 5], 1u << (group[g] & 31));
            }
            if (lane == 0) {
                keptmask_s = (int)km;
                nk_s += spec ? __popc(km) : ng;
                if (spec && __popc(km) * 4 < ng * 3) spec_s = 0;  // speculation does not pay on this problem
            }
        }
        if (spec) __syncthreads();
        LZ_MARK(7)  // keep list (+ resolve)
        if (split) cluster_wait();  // every peer is done reading its bitmap for this pass: it may be touched now
        // ---- the marked boxes leave the alive bitmap
        if (spec) {  // kill words from the kept candidates' rows: four threads per word, eight rows each
            const unsigned km = (unsigned)keptmask_s;
            for (int wb = (g0 >> 5); wb < W; wb += NT / 4) {
                const int w = wb + (tid >> 2), part = tid & 3;
                unsigned kill = 0u;
                if (w < W) {
#pragma unroll
                    for (int k = 0; k < G / 4; k++) {
                        const int g = part * (G / 4) + k;
                        if (g < ng) {
                            if ((km >> g) & 1u) kill |= sup[g * SW + w];
                            sup[g * SW + w] = 0u;
                        }
                    }
                }
                kill |= __shfl_xor_sync(FULL, kill, 1);
                kill |= __shfl_xor_sync(FULL, kill, 2);
                if (w < W && part == 0 && kill) {
                    if (split) {  // peers update the same words: atomics everywhere
                        for (int r = 0; r < C; r++) atomicAnd(cluster.map_shared_rank(&alive[w], r), ~kill);
                    } else {
                        alive[w] &= ~kill;
                    }
                }
            }
        } else {
            for (int w = (g0 >> 5) + tid; w < W; w += NT) {
                const unsigned kill = killw[w];
                if (kill) {
                    killw[w] = 0u;
                    if (split) {
                        for (int r = 0; r < C; r++) atomicAnd(cluster.map_shared_rank(&alive[w], r), ~kill);
                    } else {
                        alive[w] &= ~kill;
                    }
                }
            }
        }
#ifdef LG_LZ_TIMING
        npass++;
#endif
        if (split) cluster.sync();  // every CTA's kill words have landed everywhere before the next pass reads the bitmap
        LZ_MARK(8)  // kill words (+ cluster sync)
    }
    // all threads left the loop together.  The keep list: the kept boxes in score order (= ascending position), the first
    // max_keep of them, mapped through `order`; -1 beyond.
    if (crank == 0) {
        if (warp == 0) {
            int total = 0;
            for (int w0 = 0; w0 < W; w0 += 32) {
                const int w = w0 + lane;
                const int c = w < W ? __popc(keptw[w]) : 0;
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += v;
                }
                if (w < W) wprefix[w] = total + incl - c;
                total += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) nk_s = min(total, max_keep);
        }
        __syncthreads();
        const int nk = nk_s;
        if (tid == 0 && num_keep) num_keep[p] = nk;
        // destinations: the caller's keep row, or (gather) the packed row of this problem in every peer's buffer
        const int ndst = peers.n > 0 ? peers.n : 1;
        for (int d = 0; d < ndst; d++) {
            int64_t* const krow = peers.n > 0 ? peers.buf[d] + (peers.row0 + p) * (int64_t)keep_ld + 1 : keep + (int64_t)p * keep_ld;
            if (peers.n > 0 && tid == 0) krow[-1] = nk;
            for (int wb = 0; wb < W; wb += NT / 4) {
                const int w = wb + (tid >> 2), part = tid & 3;
                if (w < W) {
                    const unsigned word = keptw[w];
                    unsigned bits = (word >> (8 * part)) & 0xffu;
                    int off = wprefix[w] + __popc(word & ((1u << (8 * part)) - 1u));
                    while (bits && off < max_keep) {
                        const int j = w * 32 + 8 * part + __ffs(bits) - 1;
                        bits &= bits - 1;
                        krow[off++] = order ? __ldg(order + base + j) : (int64_t)j;
                    }
                }
            }
            for (int i = nk + tid; i < max_keep; i += NT) krow[i] = -1;
        }
    }
    LZ_MARK(11)  // keep list emission
    if (stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            my_tested += __shfl_xor_sync(FULL, my_tested, d);
            my_nonzero += __shfl_xor_sync(FULL, my_nonzero, d);
        }
        if (lane == 0) {
            atomicAdd(stats, (unsigned long long)my_tested);
            atomicAdd(stats + 2, (unsigned long long)my_nonzero);
        }
        if (tid == 0) atomicAdd(stats + 1, st_heavy);
#ifdef LG_LZ_TIMING
        if (tid == 0) {
            for (int ph = 0; ph < 12; ph++) atomicAdd(stats + 8 + ph, (unsigned long long)tacc[ph]);
            atomicAdd(stats + 20, (unsigned long long)npass);
            atomicMax(stats + 21, (unsigned long long)npass);
            unsigned long long tsum = 0ull;
            for (int ph = 0; ph < 12; ph++) tsum += (unsigned long long)tacc[ph];
            atomicMax(stats + 22, tsum);
        }
#endif
    }
}

// thresh < 0: every IoU, the exact zeros of disjoint boxes included, exceeds it (kernel.cu:304 `iou > thresh`), so the first box
// suppresses all the others.  The kernels above never evaluate a pair the circle test proves disjoint, so this case is answered
// directly: the best-scored box of every non-empty problem is kept, nothing else.
__global__ void __launch_bounds__(256) nms_first_only_kernel(const int64_t* __restrict__ order, const int32_t* __restrict__ counts, const int nmax,
                                                             int64_t* __restrict__ keep, const int keep_ld, const int max_keep,
                                                             int32_t* __restrict__ num_keep, const int P) {
    const int p = blockIdx.x;
    if (p >= P) return;
    const int n = problem_count(counts, p, nmax);
    const int nk = (n > 0 && max_keep > 0) ? 1 : 0;
    for (int i = threadIdx.x; i < max_keep; i += blockDim.x) {
        int64_t v = -1;
        if (i == 0 && nk) v = order ? order[(int64_t)p * nmax] : 0;
        keep[(int64_t)p * keep_ld + i] = v;
    }
    if (threadIdx.x == 0 && num_keep) num_keep[p] = nk;
}

enum { PHASE_RECORDS = 1, PHASE_MASK = 2, PHASE_SWEEP = 4, PHASE_ALL = 7 };

// per-device facts, looked up once (the C ABI keeps no state that matters: these are caches of immutable properties)
static int device_sm_count(int dev) {
    static std::atomic<int> cache[64];
    if (dev < 0 || dev >= 64) dev = 0;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

static int nms_entry(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax, float thresh, void* ws,
                     size_t ws_bytes, int64_t* keep, int keep_ld, int max_keep, int32_t* num_keep, unsigned flags, void* stream,
                     bool normal, unsigned phases = PHASE_ALL, const PeerRows* peers = nullptr) {
    if (P < 0 || nmax < 0) {
        set_error("negative size num_problems=%d nmax=%d", P, nmax);
        return LG_ERR_INVALID_ARG;
    }
    if (P == 0) return LG_OK;
    if (peers ? (nmax > 0 && !boxes) : (!num_keep || (nmax > 0 && (!boxes || !keep)))) {
        set_error("null pointer (boxes=%p keep=%p num_keep=%p)", (const void*)boxes, (void*)keep, (void*)num_keep);
        return LG_ERR_INVALID_ARG;
    }
    if (max_keep > nmax) max_keep = nmax;
    if (nmax > 0 && (max_keep < 0 || keep_ld < max_keep)) {
        set_error("max_keep=%d keep_ld=%d: need 0 <= max_keep <= keep_ld", max_keep, keep_ld);
        return LG_ERR_INVALID_ARG;
    }
    if (nmax > LG_NMS_MAX_BOXES) {
        set_error("nmax=%d exceeds LG_NMS_MAX_BOXES=%d", nmax, LG_NMS_MAX_BOXES);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (nmax == 0 && peers) {
        set_error("gather call with nmax = 0");
        return LG_ERR_INVALID_ARG;
    }
    if (nmax == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * P, st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
            return (int)e;
        }
        return LG_OK;
    }
    const size_t need = lg_nms_workspace_bytes_ex(P, nmax, normal ? 1 : 0, flags);
    if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("workspace %p of %zu B; need %zu B, 16-byte aligned", ws, ws_bytes, need);
        return LG_ERR_WORKSPACE;
    }
    const int cbk = (nmax + 63) / 64;
    const int tri = cbk * (cbk + 1) / 2;
    float4* rec = reinterpret_cast<float4*>(ws);
    const size_t rec_bytes = align_up((size_t)P * nmax * REC_F4 * sizeof(float4), 256);
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + rec_bytes);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + rec_bytes + NMS_STATS_BYTES);
    const bool strict = (flags & LG_FLAG_STRICT_FP32) != 0;
    // the lazy kernel keeps the alive bitmap and the suppression rows of its candidates in shared memory: 32 candidates per pass
    // up to ~8,000 boxes, 16 / 8 beyond; past ~50,000 boxes nothing fits next to the queues and the mask + sweep formulation
    // takes over (same keep list)
    if (!normal && thresh < 0.f && !peers) {  // (the axis-aligned kernel evaluates every pair and needs no special case)
        if (!(phases & PHASE_SWEEP)) return LG_OK;
        for (int p0 = 0; p0 < P; p0 += 65535) {
            const int pn = min(P - p0, 65535);
            nms_first_only_kernel<<<pn, 256, 0, st>>>(order ? order + (size_t)p0 * nmax : nullptr, counts ? counts + p0 : nullptr, nmax,
                                                      keep + (size_t)p0 * keep_ld, keep_ld, max_keep, num_keep + p0, pn);
        }
        return check_launch("nms_first_only_kernel");
    }
    const int gcap512 = lazy_gcap(nmax, LZ_THREADS);
    const bool full = (flags & LG_FLAG_NMS_FULL_MASK) != 0 || gcap512 == 0;
    if (peers && thresh < 0.f) {
        set_error("the fused gather does not take negative thresholds (thresh=%g): use lg_nms_batched_ex", (double)thresh);
        return LG_ERR_INVALID_ARG;
    }
    if (peers && (normal || full)) {
        set_error("the fused gather exists for the lazy rotated NMS only (nmax=%d)", nmax);
        return LG_ERR_INVALID_ARG;
    }
    int rc;
    if (!normal && !full) {
        // lazy path: records, candidate rows and the greedy resolve in ONE launch; only the rows of kept boxes are ever evaluated
        if (!(phases & PHASE_SWEEP)) return LG_OK;
        cudaError_t e = cudaMemsetAsync(stats, 0, NMS_STATS_BYTES, st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
            return (int)e;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        const int sms = device_sm_count(dev);
        // the batch goes out in slices of at most 65,535 problems (grid.x of a clustered launch is problems x cluster size)
        for (int p0 = 0; p0 < P; p0 += 32768) {
            const int pn = min(P - p0, 32768);
            // cluster size: as many SMs per problem as leaves every CTA of the batch resident at once (the kernel is bound by
            // one SM's latency per problem), and never fewer than two sweep steps per CTA
            int csize = 1;
            while (csize < LZ_MAX_CLUSTER && (int64_t)pn * csize * 2 <= sms && nmax >= csize * 2 * (LZ_THREADS / 2)) csize *= 2;
            if (flags & LG_FLAG_NMS_NO_CLUSTER) csize = 1;
            // many small problems: 256-thread CTAs, two per SM
            const bool small = csize == 1 && nmax <= 1536 && (int64_t)pn >= 2 * sms;
            const int nt = small ? 256 : LZ_THREADS;
            const int gcap = small ? lazy_gcap(nmax, 256) : gcap512;
            const LazyLayout L(nmax, nt, gcap);
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3((unsigned)(pn * csize));
            lc.blockDim = dim3((unsigned)nt);
            lc.dynamicSmemBytes = L.total;
            lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)csize;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = csize > 1 ? 1 : 0;  // a plain launch when every problem gets one CTA
            cudaError_t le;
            const float* bx = boxes + (size_t)p0 * nmax * 7;
            float4* rc4 = rec + (size_t)p0 * nmax * REC_F4;
            const int64_t* od = order ? order + (size_t)p0 * nmax : nullptr;
            const int32_t* cn = counts ? counts + p0 : nullptr;
            int64_t* kp = keep ? keep + (size_t)p0 * keep_ld : nullptr;
            int32_t* nk = num_keep ? num_keep + p0 : nullptr;
            PeerRows pr = {};
            if (peers) {
                pr = *peers;
                pr.row0 += p0;
            }
            auto launch = [&](auto kern) -> cudaError_t {
                if ((rc = set_smem(kern, L.total))) return cudaSuccess;
                return cudaLaunchKernelEx(&lc, kern, bx, rc4, od, cn, nmax, thresh, kp, keep_ld, max_keep, nk, stats, gcap, pr);
            };
            rc = 0;
            if (strict) le = csize > 1 ? launch(nms_lazy_kernel<0, true, 512>) : (small ? launch(nms_lazy_kernel<0, false, 256>) : launch(nms_lazy_kernel<0, false, 512>));
            else le = csize > 1 ? launch(nms_lazy_kernel<1, true, 512>) : (small ? launch(nms_lazy_kernel<1, false, 256>) : launch(nms_lazy_kernel<1, false, 512>));
            if (rc) return rc;
            if (le != cudaSuccess) {
                set_error("nms_lazy_kernel (cluster of %d): %s", csize, cudaGetErrorString(le));
                return (int)le;
            }
            if ((rc = check_launch("nms_lazy_kernel"))) return rc;
        }
        return LG_OK;
    }
    // mask + sweep formulation, in slices of at most 65,535 problems (grid.y)
    for (int p0 = 0; p0 < P; p0 += 65535) {
        const int pn = min(P - p0, 65535);
        const float* bx = boxes + (size_t)p0 * nmax * 7;
        float4* rc4 = rec + (size_t)p0 * nmax * REC_F4;
        const int64_t* od = order ? order + (size_t)p0 * nmax : nullptr;
        const int32_t* cn = counts ? counts + p0 : nullptr;
        unsigned long long* mk = mask + (size_t)p0 * nmax * cbk;
        dim3 mg(tri, pn);
        if (!normal) {
            if (phases & PHASE_RECORDS) {
                dim3 pg((nmax + 255) / 256, pn);
                if (strict) nms_prep_kernel<0><<<pg, 256, 0, st>>>(bx, od, cn, nmax, rc4);
                else nms_prep_kernel<1><<<pg, 256, 0, st>>>(bx, od, cn, nmax, rc4);
                if ((rc = check_launch("nms_prep_kernel"))) return rc;
            }
            if (phases & PHASE_MASK) {
                if (strict) {
                    if ((rc = set_smem(nms_mask_kernel<0>, NmsSmem::total))) return rc;
                    nms_mask_kernel<0><<<mg, NMS_THREADS, NmsSmem::total, st>>>(rc4, cn, nmax, cbk, thresh, mk);
                } else {
                    if ((rc = set_smem(nms_mask_kernel<1>, NmsSmem::total))) return rc;
                    nms_mask_kernel<1><<<mg, NMS_THREADS, NmsSmem::total, st>>>(rc4, cn, nmax, cbk, thresh, mk);
                }
                if ((rc = check_launch("nms_mask_kernel"))) return rc;
            }
        } else if (phases & PHASE_MASK) {
            if (strict) nms_normal_mask_kernel<0><<<mg, NMS_TILE, 0, st>>>(bx, od, cn, nmax, cbk, thresh, mk);
            else nms_normal_mask_kernel<1><<<mg, NMS_TILE, 0, st>>>(bx, od, cn, nmax, cbk, thresh, mk);
            if ((rc = check_launch("nms_normal_mask_kernel"))) return rc;
        }
        if (phases & PHASE_SWEEP) {
            nms_sweep_kernel<<<pn, 32, (size_t)cbk * sizeof(unsigned long long), st>>>(mk, od, cn, nmax, cbk, keep + (size_t)p0 * keep_ld, keep_ld,
                                                                                        max_keep, num_keep + p0);
            if ((rc = check_launch("nms_sweep_kernel"))) return rc;
        }
    }
    return LG_OK;
}

}  // namespace lg

extern "C" size_t lg_nms_workspace_bytes_ex(int P, int nmax, int normal, unsigned flags) {
    if (P <= 0 || nmax <= 0) return 0;
    const size_t cbk = ((size_t)nmax + 63) / 64;
    size_t b = lg::align_up((size_t)P * nmax * lg::REC_F4 * sizeof(float4), 256) + lg::NMS_STATS_BYTES + 16;
    if (normal || (flags & LG_FLAG_NMS_FULL_MASK) || lg::lazy_gcap(nmax, lg::LZ_THREADS) == 0)
        b += (size_t)P * nmax * cbk * sizeof(unsigned long long);
    return b;
}

extern "C" size_t lg_nms_workspace_bytes(int P, int nmax) { return lg_nms_workspace_bytes_ex(P, nmax, 1, LG_FLAG_NMS_FULL_MASK); }

extern "C" size_t lg_nms_stats_offset(int P, int nmax) {
    if (P <= 0 || nmax <= 0) return 0;
    return lg::align_up((size_t)P * nmax * lg::REC_F4 * sizeof(float4), 256);
}

extern "C" int lg_nms_rotated_batched(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                      float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                      unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, nmax, nmax, num_keep, flags, stream, false);
}

extern "C" int lg_nms_normal_batched(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                     float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                     unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, nmax, nmax, num_keep, flags, stream, true);
}

extern "C" int lg_nms_batched_ex(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax, float thresh,
                                 int normal, int max_keep, int64_t keep_ld, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                 unsigned flags, void* stream) {
    if (keep_ld < 0 || keep_ld > 0x7fffffffLL) {
        lg::set_error("keep_ld=%lld out of range", (long long)keep_ld);
        return LG_ERR_INVALID_ARG;
    }
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, (int)keep_ld, max_keep, num_keep, flags, stream, normal != 0);
}

extern "C" int lg_nms_rotated_gather(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax, float thresh,
                                     int max_keep, void* ws, size_t ws_bytes, int64_t* const* peer_bufs, int num_peers, int64_t row0,
                                     int32_t* num_keep, unsigned flags, void* stream) {
    if (num_peers < 1 || num_peers > LG_MAX_PEERS || !peer_bufs || row0 < 0 || max_keep < 0) {
        lg::set_error("bad gather arguments (num_peers=%d peer_bufs=%p row0=%lld max_keep=%d)", num_peers, (const void*)peer_bufs, (long long)row0, max_keep);
        return LG_ERR_INVALID_ARG;
    }
    lg::PeerRows pr = {};
    for (int r = 0; r < num_peers; r++) {
        if (!peer_bufs[r]) {
            lg::set_error("peer_bufs[%d] is NULL", r);
            return LG_ERR_INVALID_ARG;
        }
        pr.buf[r] = peer_bufs[r];
    }
    pr.n = num_peers;
    pr.row0 = row0;
    const int mk = max_keep < nmax ? max_keep : nmax;
    // the packed row pitch is 1 + max_keep as the CALLER defined it (not clamped to nmax)
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, nullptr, 1 + max_keep, mk, num_keep, flags, stream, false, lg::PHASE_ALL, &pr);
}

extern "C" int lg_nms_batched_phases(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                     float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                     unsigned flags, void* stream, int normal, unsigned phases) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, nmax, nmax, num_keep, flags, stream, normal != 0,
                         phases & lg::PHASE_ALL);
}

extern "C" int lg_nms_rotated(const float* boxes, const int64_t* order, int n, float thresh, void* ws, size_t ws_bytes,
                              int64_t* keep, int32_t* num_keep, unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, nullptr, 1, n, thresh, ws, ws_bytes, keep, n, n, num_keep, flags, stream, false);
}

extern "C" int lg_nms_normal(const float* boxes, const int64_t* order, int n, float thresh, void* ws, size_t ws_bytes,
                             int64_t* keep, int32_t* num_keep, unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, nullptr, 1, n, thresh, ws, ws_bytes, keep, n, n, num_keep, flags, stream, true);
}
