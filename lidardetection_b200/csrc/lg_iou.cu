// lg_iou.cu -- N x M rotated BEV overlap / IoU / fused 3D IoU for sm_100a.
//
// Replaces boxes_overlap_kernel / boxes_iou_bev_kernel and their launchers
// (/root/reference/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265, 378-398) and fuses the ~12
// elementwise torch kernels of boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:48-81).
//
// Design (B200: 148 SMs, 228 KB smem/SM, HBM3e):
//   1. prep kernel: one thread per box builds an 80-byte record (lg_geom.cuh) -- all trigonometry,
//      corner rotation and margin arithmetic happens N+M times instead of N*M times.
//   2. tile kernel: a CTA owns a TA x TB tile (4096 pairs) of the output.
//        cull   -- lanes run along columns; |ca - cb|^2 > (ra + rb)^2 proves the reference would
//                  return exactly +0.0; survivors are compacted (ballot + popc) into a smem queue;
//        heavy  -- the queue is drained by ALL threads, so the divergent polygon code runs at full
//                  lane occupancy whatever the survivor density (0.3 % for anchors x GT, 100 % for
//                  the dense microbench);
//        store  -- the tile is staged in smem and streamed out once, coalesced (float4, st.global.cs):
//                  DRAM traffic == algorithmic bytes 4*N*M + 28*(N+M) (+80*(N+M) record round trip).
//      Sparse workloads are therefore HBM-write bound, dense ones FP32-issue bound.
//   3. 64-bit output offsets (the reference's int32 index overflows at 2^31 pairs).
#include "lg_common.cuh"
#include "lg_geom.cuh"

namespace lg {

constexpr int IOU_THREADS = 256;
constexpr int IOU_TILE_PAIRS = 4096;

template <int FL>
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ a, int64_t n, const float* __restrict__ b,
                                                   int64_t m, float4* __restrict__ rec_a, float4* __restrict__ rec_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        make_record<FL>(a + i * 7, rec_a + i * REC_F4);
    } else if (i < n + m) {
        const int64_t j = i - n;
        make_record<FL>(b + j * 7, rec_b + j * REC_F4);
    }
}

enum { MODE_OVERLAP = 0, MODE_IOU_BEV = 1, MODE_IOU3D = 2 };

template <int TB>
struct IouSmem {
    static constexpr int TA = IOU_TILE_PAIRS / TB;
    static constexpr size_t rec_bytes = (size_t)(TA + TB) * REC_F4 * sizeof(float4);
    static constexpr size_t out_bytes = (size_t)IOU_TILE_PAIRS * sizeof(float);
    static constexpr size_t slab_bytes = (size_t)16 * IOU_THREADS * sizeof(float2);
    static constexpr size_t queue_bytes = (size_t)IOU_TILE_PAIRS * sizeof(uint16_t);
    static constexpr size_t total = rec_bytes + out_bytes + slab_bytes + queue_bytes;
};

template <int FL, int TB>
__global__ void __launch_bounds__(IOU_THREADS, 2)
    iou_tile_kernel(const float4* __restrict__ rec_a, const int64_t n, const float4* __restrict__ rec_b, const int64_t m,
                    float* __restrict__ out, const int64_t ld, const int mode, const int64_t tiles_m) {
    constexpr int TA = IOU_TILE_PAIRS / TB;
    constexpr int NT = IOU_THREADS;
    extern __shared__ float4 smem4[];
    float4* sA = smem4;
    float4* sB = sA + TA * REC_F4;
    float* sOut = reinterpret_cast<float*>(sB + TB * REC_F4);
    float2* slab = reinterpret_cast<float2*>(sOut + IOU_TILE_PAIRS);
    uint16_t* queue = reinterpret_cast<uint16_t*>(slab + 16 * NT);
    __shared__ int qcount;

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t tn = tile / tiles_m, tm = tile - tn * tiles_m;
    const int64_t row0 = tn * TA, col0 = tm * TB;
    const int na = (int)min((int64_t)TA, n - row0), nb = (int)min((int64_t)TB, m - col0);

    for (int e = tid; e < na * REC_F4; e += NT) sA[e] = __ldg(rec_a + row0 * REC_F4 + e);
    for (int e = tid; e < nb * REC_F4; e += NT) sB[e] = __ldg(rec_b + col0 * REC_F4 + e);
    for (int e = tid; e < IOU_TILE_PAIRS / 4; e += NT) reinterpret_cast<float4*>(sOut)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) qcount = 0;
    __syncthreads();

    // ---- cull: lanes along columns, rows broadcast from smem
    {
        const int col = tid % TB;
        const int lane = tid & 31;
        float4 bc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < nb) bc = sB[col * REC_F4 + 2];
#pragma unroll 4
        for (int r = tid / TB; r < TA; r += NT / TB) {
            bool surv = false;
            if (r < na && col < nb) {
                const float4 ac = sA[r * REC_F4 + 2];
                const float dx = ac.x - bc.x, dy = ac.y - bc.y, rr = ac.z + bc.z;
                surv = !(dx * dx + dy * dy > rr * rr);  // NaN => keep
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

    // ---- heavy: drain the queue with every thread
    {
        const int qn = qcount;
        for (int q = tid; q < qn; q += NT) {
            const int e = queue[q];
            const int r = e >> 8, c = e & 255;
            const float4* A = sA + r * REC_F4;
            const float4* B = sB + c * REC_F4;
            const float ov = overlap_area<FL>(A, B, slab + tid, NT);
            float res = ov;
            if (mode == MODE_IOU_BEV) res = iou_from_overlap(ov, A[2].w, B[2].w);
            else if (mode == MODE_IOU3D) res = iou3d_from_overlap(ov, A[4], B[4]);
            sOut[r * TB + c] = res;
        }
    }
    __syncthreads();

    // ---- store: one coalesced streaming pass over the tile
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
        constexpr int C4 = TB / 4;
        for (int e = tid; e < TA * C4; e += NT) {
            const int r = e / C4, c = (e % C4) * 4;
            if (r < na && c < nb) {
                const float4 v = *reinterpret_cast<const float4*>(sOut + r * TB + c);
                float* dst = out + (row0 + r) * ld + col0 + c;
                if (c + 3 < nb) {
                    __stcs(reinterpret_cast<float4*>(dst), v);
                } else {
                    __stcs(dst, v.x);
                    if (c + 1 < nb) __stcs(dst + 1, v.y);
                    if (c + 2 < nb) __stcs(dst + 2, v.z);
                }
            }
        }
    } else {
        for (int e = tid; e < TA * TB; e += NT) {
            const int r = e / TB, c = e % TB;
            if (r < na && c < nb) __stcs(out + (row0 + r) * ld + col0 + c, sOut[e]);
        }
    }
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

template <int FL, int TB>
static int launch_tiles(const float4* ra, int64_t n, const float4* rb, int64_t m, float* out, int64_t ld, int mode,
                        cudaStream_t st) {
    constexpr int TA = IOU_TILE_PAIRS / TB;
    const int64_t tiles_n = (n + TA - 1) / TA, tiles_m = (m + TB - 1) / TB;
    const int64_t tiles = tiles_n * tiles_m;
    if (tiles > 0x7fffffffLL) {
        set_error("%lld tiles exceed the 1-D grid limit; split the call by row blocks", (long long)tiles);
        return LG_ERR_TOO_LARGE;
    }
    auto kern = iou_tile_kernel<FL, TB>;
    int rc = set_smem(kern, IouSmem<TB>::total);
    if (rc) return rc;
    kern<<<(unsigned)tiles, IOU_THREADS, IouSmem<TB>::total, st>>>(ra, n, rb, m, out, ld, mode, tiles_m);
    return check_launch("iou_tile_kernel");
}

template <int FL>
static int run_iou(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws, int mode,
                   cudaStream_t st) {
    float4* ra = reinterpret_cast<float4*>(ws);
    float4* rb = ra + n * REC_F4;
    const int64_t total = n + m;
    prep_kernel<FL><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, n, b, m, ra, rb);
    int rc = check_launch("prep_kernel");
    if (rc) return rc;
    if (m <= 32) return launch_tiles<FL, 32>(ra, n, rb, m, out, ld, mode, st);
    if (m <= 64) return launch_tiles<FL, 64>(ra, n, rb, m, out, ld, mode, st);
    return launch_tiles<FL, 128>(ra, n, rb, m, out, ld, mode, st);
}

static int iou_entry(const float* a, int64_t n, const float* b, int64_t m, float* out, int64_t ld, void* ws, size_t ws_bytes,
                     unsigned flags, void* stream, int mode) {
    int rc = check_args(a, n, b, m, out, ld, ws, ws_bytes);
    if (rc || n == 0 || m == 0) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (flags & LG_FLAG_STRICT_FP32) return run_iou<0>(a, n, b, m, out, ld, ws, mode, st);
    return run_iou<1>(a, n, b, m, out, ld, ws, mode, st);
}

}  // namespace lg

extern "C" size_t lg_iou_workspace_bytes(int64_t n, int64_t m) {
    if (n < 0 || m < 0) return 0;
    return (size_t)(n + m) * lg::REC_F4 * sizeof(float4) + 16;
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
