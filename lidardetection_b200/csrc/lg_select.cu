// lg_select.cu -- SURVEY 8f-1: the selection steps around NMS in the post-processing front end
// (pcdet/models/model_utils/model_nms_utils.py:6-25 per frame: score mask -> torch.topk -> gather -> nms -> truncate -> map back;
//  the reference runs that chain once per frame / per class from a Python loop, detector3d_template.py:190-260).
//
//   lg_select_topk    score threshold + sorted top-k + box gather for P problems in one launch (one 1024-thread CTA per
//                     problem): replaces  where / topk / count / gather  of the batched front end.
//   lg_select_finish  truncate the keep lists to NMS_POST_MAXSIZE and map them back to candidate indices and scores.
//
// Selection rule: descending score, ties broken by ASCENDING candidate index (torch.topk leaves the order of equal scores
// unspecified; this rule is deterministic).  A NaN score never passes a threshold (`score >= thresh` is false, as in the
// reference's mask); without a threshold a positive NaN sorts above +inf, as torch.topk orders it.
//
// Structure per problem: one streaming pass turns every candidate that passes the threshold into a 64-bit composite
// (order-preserving score bits << 32 | ~index) in the workspace.  If no more than k pass -- the normal case with SECOND's
// SCORE_THRESH -- they are all selected; otherwise an MSB-first radix select over the composites (8 bits per pass, stops as
// soon as a digit group is taken whole) finds the k-th composite and a collect pass keeps the ones at or above it.  The
// selected composites (<= 4096, kept in ascending candidate index by ordered block-wide compaction) are sorted by a stable
// 4-pass LSD radix sort on the score key in shared memory and written out with their boxes.
#include "lg_common.cuh"

namespace lg {
namespace sel {

constexpr int NT = 1024;
constexpr int KMAX = 4096;

__device__ __forceinline__ unsigned long long compose(float v, unsigned idx) {
    unsigned b = __float_as_uint(v);
    if (b == 0x80000000u) b = 0u;          // -0.0 == +0.0, as every comparison-based sort (torch.sort / topk) sees them
    if (v != v) b = 0x7FFFFFFFu;           // any NaN, whatever its sign bit, ranks above +inf (torch's order)
    const unsigned key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // larger float -> larger unsigned
    return ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

constexpr int NW = NT / 32;
constexpr size_t SMEM_BYTES = 2 * (size_t)KMAX * sizeof(unsigned long long) + (size_t)NW * 256 * sizeof(uint16_t);  // 80 KB

// Block-wide ORDERED append: thread t's element comes before thread t + 1's.  Returns the slot of this thread's element
// (meaningful where pred) and advances `running` (a per-thread copy of the block-uniform count) by the block's total.
// Contains two barriers; s_warp is 32 ints of shared scratch.
__device__ __forceinline__ int ordered_append(const bool pred, int& running, int* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    const int c = s_warp[lane];  // NW == 32: one warp count per lane
    int before = lane < warp ? c : 0, total = c;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, d);
        total += __shfl_xor_sync(0xffffffffu, total, d);
    }
    const int slot = running + before + __popc(m & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return slot;
}

// Stable LSD radix sort (4 passes of 8 bits) of KMAX composites by DESCENDING score key (the high 32 bits); equal keys keep
// their input order, which is ascending candidate index.  Warp w ranks elements [128 w, 128 w + 128) in position order with
// match.any (striped: lane l, slice e <-> position 128 w + 32 e + l); per-warp digit counters are scanned (digit-major) into
// scatter offsets.  Result in bufA.
// E = elements per thread (1, 2 or 4): the first 1024 E slots are sorted (warp w ranks [32 E w, 32 E (w + 1))).
__device__ __forceinline__ void radix_sort_desc(unsigned long long* bufA, unsigned long long* bufB, uint16_t* cnt, int* s_warp, const int E) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned long long *src = bufA, *dst = bufB;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        for (int i = tid; i < NW * 256 / 2; i += NT) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
        __syncthreads();
        unsigned long long v[4];
        int r[4], dg[4];
        uint16_t* wc = cnt + warp * 256;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (e < E) {  // block-uniform
                v[e] = src[32 * E * warp + 32 * e + lane];
                const int d = 255 - (int)((v[e] >> shift) & 255ull);  // larger key -> lower bin
                dg[e] = d;
                const unsigned m = __match_any_sync(0xffffffffu, d);
                const int base = wc[d];
                __syncwarp();
                if (lane == __ffs(m) - 1) wc[d] = (uint16_t)(base + __popc(m));
                __syncwarp();
                r[e] = base + __popc(m & lt);
            }
        }
        __syncthreads();
        {   // exclusive scan of cnt in (digit, warp) order: thread t owns digit t >> 2, warps 8 (t & 3) .. + 7
            const int d = tid >> 2, w0 = 8 * (tid & 3);
            int c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                c[k] = cnt[(w0 + k) * 256 + d];
                sum += c[k];
            }
            int incl = sum;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, dd);
                if (lane >= dd) incl += o;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            const int wt = s_warp[lane];
            int wbefore = lane < warp ? wt : 0;
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) wbefore += __shfl_xor_sync(0xffffffffu, wbefore, dd);
            int run = wbefore + incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                cnt[(w0 + k) * 256 + d] = (uint16_t)run;
                run += c[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (e < E) dst[wc[dg[e]] + r[e]] = v[e];
        __syncthreads();
        unsigned long long* t = src;
        src = dst;
        dst = t;
    }
}

__global__ void __launch_bounds__(NT) select_topk_kernel(const float* __restrict__ scores, const int64_t n, const int k, const float thresh,
                                                          const int use_thresh, const float* __restrict__ boxes,
                                                          const int64_t box_frame_stride, const int64_t box_row_stride,
                                                          const int problems_per_frame, unsigned long long* __restrict__ ws,
                                                          int64_t* __restrict__ top_idx, int32_t* __restrict__ counts,
                                                          float* __restrict__ top_boxes) {
    extern __shared__ unsigned long long smem_sel[];
    unsigned long long* s_sel = smem_sel;          // KMAX composites
    unsigned long long* s_alt = smem_sel + KMAX;   // the radix sort's second buffer
    uint16_t* s_cnt = reinterpret_cast<uint16_t*>(smem_sel + 2 * KMAX);
    __shared__ int s_hist[256];
    __shared__ int s_warp[NW];
    __shared__ int s_digit, s_want, s_stop;
    const int p = blockIdx.x, tid = threadIdx.x;
    const float* s = scores + (int64_t)p * n;
    unsigned long long* cand = ws + (int64_t)p * n;
    // ---- pass A: candidates that pass the threshold -> composites, in ascending candidate index
    int V = 0;
    for (int64_t i0 = 0; i0 < n; i0 += NT) {
        const int64_t i = i0 + tid;
        const float v = i < n ? __ldg(s + i) : 0.f;
        const bool ok = i < n && (use_thresh ? (v >= thresh) : true);
        const int slot = ordered_append(ok, V, s_warp);
        if (ok) cand[slot] = compose(v, (unsigned)i);
    }
    __syncthreads();
    const int cnt = V < k ? V : k;
    if (V <= k) {
        for (int i = tid; i < V; i += NT) s_sel[i] = cand[i];
    } else {
        // ---- radix select: the k-th largest composite.  prefix/mask describe the digit group still undecided.
        unsigned long long prefix = 0ull, mask = 0ull, bound = 0ull;
        int want = k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int d = tid; d < 256; d += NT) s_hist[d] = 0;
            __syncthreads();
            for (int i = tid; i < V; i += NT) {
                const unsigned long long c = cand[i];
                if ((c & mask) == prefix) {
                    const int d = (int)((c >> shift) & 255ull);
                    const unsigned act = __activemask();
                    const unsigned same = __match_any_sync(act, d);
                    if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&s_hist[d], __popc(same));
                }
            }
            __syncthreads();
            if (tid == 0) {
                int above = 0, d = 255;
                for (; d > 0; --d) {
                    if (above + s_hist[d] >= want) break;
                    above += s_hist[d];
                }
                s_digit = d;
                s_want = want - above;                       // still to take inside digit group d
                s_stop = (s_hist[d] == want - above) ? 1 : 0;  // the whole group is taken: every lower bit is free
            }
            __syncthreads();
            prefix |= (unsigned long long)s_digit << shift;
            mask |= 255ull << shift;
            want = s_want;
            bound = prefix;
            const int stop = s_stop;
            __syncthreads();
            if (stop) break;
        }
        // composites are distinct, so exactly k of them are >= bound; collected in ascending candidate index
        int taken = 0;
        for (int i0 = 0; i0 < V; i0 += NT) {
            const int i = i0 + tid;
            const unsigned long long c = i < V ? cand[i] : 0ull;
            const bool ok = i < V && c >= bound;
            const int slot = ordered_append(ok, taken, s_warp);
            if (ok && slot < KMAX) s_sel[slot] = c;
        }
    }
    // ---- sort the first 1024 E >= cnt slots (padding = 0 sorts last): descending score, ties in input order = ascending index
    const int E = cnt <= NT ? 1 : (cnt <= 2 * NT ? 2 : 4);
    for (int i = cnt + tid; i < NT * E; i += NT) s_sel[i] = 0ull;
    __syncthreads();
    radix_sort_desc(s_sel, s_alt, s_cnt, s_warp, E);
    // ---- outputs
    if (tid == 0) counts[p] = cnt;
    const float* fb = boxes ? boxes + (int64_t)(p / problems_per_frame) * box_frame_stride : nullptr;
    for (int i = tid; i < k; i += NT) {
        const int64_t idx = i < cnt ? (int64_t)(0xFFFFFFFFu - (unsigned)(s_sel[i] & 0xFFFFFFFFull)) : 0;
        top_idx[(int64_t)p * k + i] = idx;
    }
    if (top_boxes) {
        for (int e = tid; e < k * 7; e += NT) {
            const int i = e / 7, c = e - i * 7;
            float v = 0.f;
            if (i < cnt) {
                const int64_t idx = (int64_t)(0xFFFFFFFFu - (unsigned)(s_sel[i] & 0xFFFFFFFFull));
                v = __ldg(fb + idx * box_row_stride + c);
            }
            top_boxes[((int64_t)p * k) * 7 + e] = v;
        }
    }
}

// selected[p, j] = top_idx[p, keep[p, j]] for j < min(num_keep[p], post), else -1; scores alike (0 beyond)
__global__ void __launch_bounds__(256) select_finish_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ num_keep,
                                                            const int64_t* __restrict__ top_idx, const float* __restrict__ scores,
                                                            const int P, const int64_t n, const int k, const int post,
                                                            int64_t* __restrict__ selected, int32_t* __restrict__ num_out,
                                                            float* __restrict__ sel_scores) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)P * post) return;
    const int p = (int)(e / post), j = (int)(e - (int64_t)p * post);
    const int m = min(num_keep[p], post);
    if (j == 0) num_out[p] = m;
    int64_t idx = -1;
    float sc = 0.f;
    if (j < m && j < k) {
        idx = top_idx[(int64_t)p * k + keep[(int64_t)p * k + j]];
        sc = scores[(int64_t)p * n + idx];
    }
    selected[e] = idx;
    sel_scores[e] = sc;
}

}  // namespace sel
}  // namespace lg

extern "C" size_t lg_select_workspace_bytes(int num_problems, int64_t n) {
    if (num_problems < 0 || n < 0) return 0;
    return (size_t)num_problems * (size_t)n * sizeof(unsigned long long);
}

extern "C" int lg_select_topk(const float* scores, int num_problems, int64_t n, int k, float score_thresh, int use_thresh,
                              const float* boxes, int64_t box_frame_stride, int64_t box_row_stride, int problems_per_frame,
                              int64_t* top_idx, int32_t* counts, float* top_boxes, void* ws, size_t ws_bytes, unsigned flags, void* stream) {
    using namespace lg;
    (void)flags;
    if (num_problems < 0 || n < 0 || k < 0 || problems_per_frame < 1) {
        set_error("invalid size (num_problems=%d n=%lld k=%d problems_per_frame=%d)", num_problems, (long long)n, k, problems_per_frame);
        return LG_ERR_INVALID_ARG;
    }
    if (num_problems == 0 || k == 0) return LG_OK;
    if (k > sel::KMAX || n > 0x7FFFFFFFLL) {
        set_error("k=%d exceeds LG_SELECT_MAX_K=%d (or n=%lld exceeds 2^31 - 1)", k, sel::KMAX, (long long)n);
        return LG_ERR_TOO_LARGE;
    }
    if (!scores || !top_idx || !counts || (top_boxes && !boxes)) {
        set_error("null pointer (scores=%p top_idx=%p counts=%p boxes=%p)", (const void*)scores, (void*)top_idx, (void*)counts, (const void*)boxes);
        return LG_ERR_INVALID_ARG;
    }
    if (!ws || ws_bytes < lg_select_workspace_bytes(num_problems, n)) {
        set_error("workspace too small: need %zu bytes, got %zu", lg_select_workspace_bytes(num_problems, n), ws_bytes);
        return LG_ERR_WORKSPACE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = set_smem(sel::select_topk_kernel, sel::SMEM_BYTES);
    if (rc) return rc;
    sel::select_topk_kernel<<<(unsigned)num_problems, sel::NT, sel::SMEM_BYTES, st>>>(scores, n, k, score_thresh, use_thresh, boxes, box_frame_stride,
                                                                        box_row_stride, problems_per_frame,
                                                                        static_cast<unsigned long long*>(ws), top_idx, counts, top_boxes);
    return check_launch("select_topk_kernel");
}

extern "C" int lg_select_finish(const int64_t* keep, const int32_t* num_keep, const int64_t* top_idx, const float* scores, int num_problems,
                                int64_t n, int k, int post, int64_t* selected, int32_t* num_out, float* sel_scores, void* stream) {
    using namespace lg;
    if (num_problems < 0 || n < 0 || k < 0 || post < 0) {
        set_error("negative size (num_problems=%d n=%lld k=%d post=%d)", num_problems, (long long)n, k, post);
        return LG_ERR_INVALID_ARG;
    }
    if (num_problems == 0 || post == 0) return LG_OK;
    if (!keep || !num_keep || !top_idx || !scores || !selected || !num_out || !sel_scores) {
        set_error("null pointer");
        return LG_ERR_INVALID_ARG;
    }
    const int64_t total = (int64_t)num_problems * post;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    sel::select_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(keep, num_keep, top_idx, scores, num_problems, n, k, post, selected,
                                                                               num_out, sel_scores);
    return check_launch("select_finish_kernel");
}
