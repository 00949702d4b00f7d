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
#include "lg_common.cuh"
#include "lg_geom.cuh"

namespace lg {

constexpr int NMS_THREADS = 256;
constexpr int NMS_TILE = 64;

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
    static constexpr size_t slab_bytes = (size_t)16 * NMS_THREADS * sizeof(float2);
    static constexpr size_t queue_bytes = (size_t)NMS_TILE * NMS_TILE * sizeof(uint16_t);
    static constexpr size_t mask_bytes = (size_t)NMS_TILE * sizeof(unsigned long long);
    static constexpr size_t total = rec_bytes + slab_bytes + queue_bytes + mask_bytes;
};

template <int FL>
__global__ void __launch_bounds__(NMS_THREADS, 2)
    nms_mask_kernel(const float4* __restrict__ rec, const int32_t* __restrict__ counts, const int nmax, const int cbk,
                    const float thresh, unsigned long long* __restrict__ mask) {
    constexpr int NT = NMS_THREADS, T = NMS_TILE;
    extern __shared__ float4 smem4[];
    float4* sA = smem4;
    float4* sB = sA + T * REC_F4;
    float2* slab = reinterpret_cast<float2*>(sB + T * REC_F4);
    uint16_t* queue = reinterpret_cast<uint16_t*>(slab + 16 * NT);
    unsigned int* smask = reinterpret_cast<unsigned int*>(queue + T * T);  // 64 x (lo, hi)
    __shared__ int qcount;

    const int p = blockIdx.y;
    const int n = problem_count(counts, p, nmax);
    const int nb = (n + T - 1) / T;
    int rb, cb;
    tri_decode(blockIdx.x, cbk, rb, cb);
    if (rb >= nb || cb >= nb) return;  // tile outside this problem's boxes

    const int tid = threadIdx.x;
    const int na = min(T, n - rb * T), ncol = min(T, n - cb * T);
    const float4* gA = rec + ((int64_t)p * nmax + (int64_t)rb * T) * REC_F4;
    const float4* gB = rec + ((int64_t)p * nmax + (int64_t)cb * T) * REC_F4;
    for (int e = tid; e < na * REC_F4; e += NT) sA[e] = __ldg(gA + e);
    for (int e = tid; e < ncol * REC_F4; e += NT) sB[e] = __ldg(gB + e);
    if (tid < 2 * T) smask[tid] = 0u;
    if (tid == 0) qcount = 0;
    __syncthreads();

    {
        const int col = tid % T, lane = tid & 31;
        const bool diag = (rb == cb);
        float4 bc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < ncol) bc = sB[col * REC_F4 + 2];
#pragma unroll 4
        for (int r = tid / T; r < T; r += NT / T) {
            bool surv = false;
            if (r < na && col < ncol && (!diag || col > r)) {
                const float4 ac = sA[r * REC_F4 + 2];
                const float dx = ac.x - bc.x, dy = ac.y - bc.y, rr = ac.z + bc.z;
                surv = !(dx * dx + dy * dy > rr * rr);
            }
            const unsigned msk = __ballot_sync(0xffffffffu, surv);
            if (msk) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&qcount, __popc(msk));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (surv) queue[base + __popc(msk & ((1u << lane) - 1u))] = (uint16_t)((r << 8) | col);
            }
        }
    }
    __syncthreads();

    {
        const int qn = qcount;
        for (int q = tid; q < qn; q += NT) {
            const int e = queue[q];
            const int r = e >> 8, c = e & 255;
            const float4* A = sA + r * REC_F4;  // row = the higher-scoring box: iou_bev(row, col), kernel.cu:304
            const float4* B = sB + c * REC_F4;
            const float ov = overlap_area<FL>(A, B, slab + tid, NT);
            const float iou = iou_from_overlap(ov, A[2].w, B[2].w);
            if (iou > thresh) atomicOr(&smask[2 * r + (c >> 5)], 1u << (c & 31));
        }
    }
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
                     const int32_t* __restrict__ counts, const int nmax, const int cbk, int64_t* __restrict__ keep,
                     int32_t* __restrict__ num_keep) {
    extern __shared__ unsigned long long remv[];
    const int p = blockIdx.x, lane = threadIdx.x;
    const int n = problem_count(counts, p, nmax);
    const int nb = (n + 63) / 64;
    const int64_t base = (int64_t)p * nmax;
    const unsigned long long* M = mask + base * cbk;
    for (int w = lane; w < nb; w += 32) remv[w] = 0ull;
    __syncwarp();
    int nk = 0;
    for (int b = 0; b < nb; b++) {
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
            if ((kept >> lane) & 1ull) keep[base + nk + __popcll(below0)] = order ? order[base + r0] : (int64_t)r0;
            if ((kept >> (lane + 32)) & 1ull) keep[base + nk + __popcll(below1)] = order ? order[base + r1] : (int64_t)r1;
        }
        nk += __popcll(kept);
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
    for (int i = nk + lane; i < nmax; i += 32) keep[base + i] = -1;
}

enum { PHASE_RECORDS = 1, PHASE_MASK = 2, PHASE_SWEEP = 4, PHASE_ALL = 7 };

static int nms_entry(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax, float thresh, void* ws,
                     size_t ws_bytes, int64_t* keep, int32_t* num_keep, unsigned flags, void* stream, bool normal,
                     unsigned phases = PHASE_ALL) {
    if (P < 0 || nmax < 0) {
        set_error("negative size num_problems=%d nmax=%d", P, nmax);
        return LG_ERR_INVALID_ARG;
    }
    if (P == 0) return LG_OK;
    if (!num_keep || (nmax > 0 && (!boxes || !keep))) {
        set_error("null pointer (boxes=%p keep=%p num_keep=%p)", (const void*)boxes, (void*)keep, (void*)num_keep);
        return LG_ERR_INVALID_ARG;
    }
    if (nmax > LG_NMS_MAX_BOXES) {
        set_error("nmax=%d exceeds LG_NMS_MAX_BOXES=%d", nmax, LG_NMS_MAX_BOXES);
        return LG_ERR_TOO_LARGE;
    }
    if (P > 65535) {
        set_error("num_problems=%d exceeds 65535; split the batch", P);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (nmax == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * P, st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
            return (int)e;
        }
        return LG_OK;
    }
    const size_t need = lg_nms_workspace_bytes(P, nmax);
    if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("workspace %p of %zu B; need %zu B, 16-byte aligned", ws, ws_bytes, need);
        return LG_ERR_WORKSPACE;
    }
    const int cbk = (nmax + 63) / 64;
    const int tri = cbk * (cbk + 1) / 2;
    float4* rec = reinterpret_cast<float4*>(ws);
    const size_t rec_bytes = align_up((size_t)P * nmax * REC_F4 * sizeof(float4), 256);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + rec_bytes);
    const bool strict = (flags & LG_FLAG_STRICT_FP32) != 0;
    int rc;
    if (!normal) {
        if (phases & PHASE_RECORDS) {
            dim3 pg((nmax + 255) / 256, P);
            if (strict) nms_prep_kernel<0><<<pg, 256, 0, st>>>(boxes, order, counts, nmax, rec);
            else nms_prep_kernel<1><<<pg, 256, 0, st>>>(boxes, order, counts, nmax, rec);
            if ((rc = check_launch("nms_prep_kernel"))) return rc;
        }
        if (!(phases & PHASE_MASK)) goto sweep;
        dim3 mg(tri, P);
        if (strict) {
            if ((rc = set_smem(nms_mask_kernel<0>, NmsSmem::total))) return rc;
            nms_mask_kernel<0><<<mg, NMS_THREADS, NmsSmem::total, st>>>(rec, counts, nmax, cbk, thresh, mask);
        } else {
            if ((rc = set_smem(nms_mask_kernel<1>, NmsSmem::total))) return rc;
            nms_mask_kernel<1><<<mg, NMS_THREADS, NmsSmem::total, st>>>(rec, counts, nmax, cbk, thresh, mask);
        }
        if ((rc = check_launch("nms_mask_kernel"))) return rc;
    } else if (phases & PHASE_MASK) {
        dim3 mg(tri, P);
        if (strict) nms_normal_mask_kernel<0><<<mg, NMS_TILE, 0, st>>>(boxes, order, counts, nmax, cbk, thresh, mask);
        else nms_normal_mask_kernel<1><<<mg, NMS_TILE, 0, st>>>(boxes, order, counts, nmax, cbk, thresh, mask);
        if ((rc = check_launch("nms_normal_mask_kernel"))) return rc;
    }
sweep:
    if (!(phases & PHASE_SWEEP)) return LG_OK;
    nms_sweep_kernel<<<P, 32, (size_t)cbk * sizeof(unsigned long long), st>>>(mask, order, counts, nmax, cbk, keep, num_keep);
    return check_launch("nms_sweep_kernel");
}

}  // namespace lg

extern "C" size_t lg_nms_workspace_bytes(int P, int nmax) {
    if (P <= 0 || nmax <= 0) return 0;
    const size_t cbk = ((size_t)nmax + 63) / 64;
    return lg::align_up((size_t)P * nmax * lg::REC_F4 * sizeof(float4), 256) + (size_t)P * nmax * cbk * sizeof(unsigned long long) + 16;
}

extern "C" int lg_nms_rotated_batched(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                      float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                      unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, num_keep, flags, stream, false);
}

extern "C" int lg_nms_normal_batched(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                     float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                     unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, num_keep, flags, stream, true);
}

extern "C" int lg_nms_batched_phases(const float* boxes, const int64_t* order, const int32_t* counts, int P, int nmax,
                                     float thresh, void* ws, size_t ws_bytes, int64_t* keep, int32_t* num_keep,
                                     unsigned flags, void* stream, int normal, unsigned phases) {
    return lg::nms_entry(boxes, order, counts, P, nmax, thresh, ws, ws_bytes, keep, num_keep, flags, stream, normal != 0,
                         phases & lg::PHASE_ALL);
}

extern "C" int lg_nms_rotated(const float* boxes, const int64_t* order, int n, float thresh, void* ws, size_t ws_bytes,
                              int64_t* keep, int32_t* num_keep, unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, nullptr, 1, n, thresh, ws, ws_bytes, keep, num_keep, flags, stream, false);
}

extern "C" int lg_nms_normal(const float* boxes, const int64_t* order, int n, float thresh, void* ws, size_t ws_bytes,
                             int64_t* keep, int32_t* num_keep, unsigned flags, void* stream) {
    return lg::nms_entry(boxes, order, nullptr, 1, n, thresh, ws, ws_bytes, keep, num_keep, flags, stream, true);
}
