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
// selected composites (<= 4096) are sorted by a bitonic network (registers / warp shuffles / shared memory) and written
// out with their boxes.
#include "lg_common.cuh"

namespace lg {
namespace sel {

constexpr int NT = 1024;
constexpr int KMAX = 4096;

__device__ __forceinline__ unsigned long long compose(float v, unsigned idx) {
    const unsigned b = __float_as_uint(v);
    const unsigned key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // larger float -> larger unsigned
    return ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

// warp-aggregated append: returns this lane's slot (valid only where pred), advancing *counter by the warp's count
__device__ __forceinline__ int warp_append(bool pred, int* counter) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(NT) select_topk_kernel(const float* __restrict__ scores, const int64_t n, const int k, const float thresh,
                                                          const int use_thresh, const float* __restrict__ boxes,
                                                          const int64_t box_frame_stride, const int64_t box_row_stride,
                                                          const int problems_per_frame, unsigned long long* __restrict__ ws,
                                                          int64_t* __restrict__ top_idx, int32_t* __restrict__ counts,
                                                          float* __restrict__ top_boxes) {
    __shared__ unsigned long long s_sel[KMAX];
    __shared__ int s_hist[256];
    __shared__ int s_cnt, s_digit, s_want, s_stop;
    const int p = blockIdx.x, tid = threadIdx.x;
    const float* s = scores + (int64_t)p * n;
    unsigned long long* cand = ws + (int64_t)p * n;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // ---- pass A: candidates that pass the threshold -> composites (any order)
    for (int64_t i0 = 0; i0 < n; i0 += NT) {
        const int64_t i = i0 + tid;
        const float v = i < n ? __ldg(s + i) : 0.f;
        const bool ok = i < n && (use_thresh ? (v >= thresh) : true);
        const int slot = warp_append(ok, &s_cnt);
        if (ok) cand[slot] = compose(v, (unsigned)i);
    }
    __syncthreads();
    const int V = s_cnt;
    const int cnt = V < k ? V : k;
    __syncthreads();
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
        // composites are distinct, so exactly k of them are >= bound
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int i0 = 0; i0 < V; i0 += NT) {
            const int i = i0 + tid;
            const unsigned long long c = i < V ? cand[i] : 0ull;
            const bool ok = i < V && c >= bound;
            const int slot = warp_append(ok, &s_cnt);
            if (ok && slot < KMAX) s_sel[slot] = c;
        }
    }
    __syncthreads();
    // ---- bitonic sort, descending, of all KMAX slots (padding = 0 sorts last).  Every thread owns 4 consecutive elements in
    // registers: exchange distances 1, 2 stay inside the thread, 4..64 go through warp shuffles, only distances >= 128 cross
    // warps through shared memory (15 of the 78 stages need a CTA barrier)
    for (int i = cnt + tid; i < KMAX; i += NT) s_sel[i] = 0ull;
    __syncthreads();
    {
        unsigned long long v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = s_sel[4 * tid + e];
        auto keep = [](unsigned long long own, unsigned long long other, bool want_max) {
            return (own > other) == want_max ? own : other;
        };
        for (int kk = 2; kk <= KMAX; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                if (j >= 128) {
                    __syncthreads();  // the previous readers of s_sel are done
#pragma unroll
                    for (int e = 0; e < 4; e++) s_sel[4 * tid + e] = v[e];
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int x = 4 * tid + e;
                        const unsigned long long o = s_sel[x ^ j];
                        v[e] = keep(v[e], o, ((x & j) == 0) == ((x & kk) == 0));
                    }
                } else if (j >= 4) {
                    const int d = j >> 2;  // partner thread inside the warp
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int x = 4 * tid + e;
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[e], d);
                        v[e] = keep(v[e], o, ((x & j) == 0) == ((x & kk) == 0));
                    }
                } else {
                    // x = 4 tid + e: bits 0, 1 of x are e's, and kk >= 4 here unless kk == 2 (then (x & kk) tests bit 1 of e)
                    unsigned long long w[4];
                    if (j == 2) {
#pragma unroll
                        for (int e = 0; e < 4; e++) w[e] = keep(v[e], v[e ^ 2], ((e & 2) == 0) == (((4 * tid + e) & kk) == 0));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++) w[e] = keep(v[e], v[e ^ 1], ((e & 1) == 0) == (((4 * tid + e) & kk) == 0));
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) v[e] = w[e];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; e++) s_sel[4 * tid + e] = v[e];
    }
    __syncthreads();
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
    sel::select_topk_kernel<<<(unsigned)num_problems, sel::NT, 0, st>>>(scores, n, k, score_thresh, use_thresh, boxes, box_frame_stride,
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
