// lg_points.cu -- points-in-boxes for sm_100a.
//
// Replaces points_in_boxes_kernel + launcher (/root/reference/pcdet/ops/roiaware_pool3d/src/
// roiaware_pool3d_kernel.cu:16-36, 313-359) and offers the all-pairs mask form of points_in_boxes_cpu
// (roiaware_pool3d.cpp:121-168) on the GPU.
//
// Predicate (check_pt_in_box3d), reproduced bit-exactly:
//     |z - cz| <= dz/2                      (closed; the reference compares in double, which for
//                                            float operands equals the float compare against dz*0.5f)
//     |lx| < dx/2 + MARGIN, |ly| < dy/2 + MARGIN   (open; compared in DOUBLE in the reference).
//       For a float v and a double D,  (double)v < D  <=>  v < RU(D)  with RU = round-up to float,
//       so the per-box thresholds are converted once (cvt.rp.f32.f64) and the per-point compare is FP32.
//     lx = fma(sx, cosa, -(sy*sina)),  ly = fma(sy, cosa, sx*sina)    with cosa = cosf(-rz), sina = sinf(-rz)
//       -- the contraction ptxas applies to lidar_to_local_coords on sm_100a (FL = 1); FL = 0 is the
//       un-contracted CPU build.
// Per-box trigonometry and thresholds are hoisted into a 32-byte record kept in shared memory.
#include "lg_common.cuh"
#include "lg_geom.cuh"

namespace lg {

constexpr int PIB_THREADS = 256;
constexpr int PIB_PPT = 4;  // points per thread

// record: (cx, cy, cz, dz/2) (cosa, sina, tx, ty)
__device__ __forceinline__ void make_pib_record(const float* __restrict__ box, const float margin, float4& r0, float4& r1) {
    const float cx = box[0], cy = box[1], cz = box[2], dx = box[3], dy = box[4], dz = box[5], rz = box[6];
    const float cosa = cosf(-rz), sina = sinf(-rz);
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

template <int FL>
__global__ void __launch_bounds__(PIB_THREADS)
    pib_idx_kernel(const float* __restrict__ boxes, const float* __restrict__ pts, int32_t* __restrict__ out, const int T,
                   const int64_t M) {
    extern __shared__ float4 srec[];  // 2 * T
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const float* fb = boxes + (int64_t)b * T * 7;
    for (int k = tid; k < T; k += PIB_THREADS) {
        float4 r0, r1;
        make_pib_record(fb + k * 7, 1e-5f, r0, r1);
        srec[2 * k] = r0;
        srec[2 * k + 1] = r1;
    }
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * (PIB_THREADS * PIB_PPT) + tid;
    const float* fp = pts + (int64_t)b * M * 3;
    float x[PIB_PPT], y[PIB_PPT], z[PIB_PPT];
    int res[PIB_PPT];
#pragma unroll
    for (int u = 0; u < PIB_PPT; u++) {
        const int64_t p = p0 + (int64_t)u * PIB_THREADS;
        res[u] = -1;
        x[u] = y[u] = z[u] = 0.f;
        if (p < M) {
            x[u] = __ldg(fp + p * 3);
            y[u] = __ldg(fp + p * 3 + 1);
            z[u] = __ldg(fp + p * 3 + 2);
        }
    }
    for (int k = 0; k < T; k++) {
        const float4 r0 = srec[2 * k], r1 = srec[2 * k + 1];
#pragma unroll
        for (int u = 0; u < PIB_PPT; u++)
            if (res[u] < 0 && pt_in_box<FL>(x[u], y[u], z[u], r0, r1)) res[u] = k;  // lowest index wins
    }
#pragma unroll
    for (int u = 0; u < PIB_PPT; u++) {
        const int64_t p = p0 + (int64_t)u * PIB_THREADS;
        if (p < M) out[(int64_t)b * M + p] = res[u];
    }
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
        make_pib_record(boxes + (box0 + tid) * 7, margin, r0, r1);
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
    if (B > 65535) {
        set_error("batch=%d exceeds 65535; split the batch", B);
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)T * 2 * sizeof(float4);
    dim3 grid((unsigned)((M + PIB_THREADS * PIB_PPT - 1) / (PIB_THREADS * PIB_PPT)), B);
    int rc;
    if (flags & LG_FLAG_STRICT_FP32) {
        if ((rc = set_smem(pib_idx_kernel<0>, smem))) return rc;
        pib_idx_kernel<0><<<grid, PIB_THREADS, smem, st>>>(boxes, pts, out, T, M);
    } else {
        if ((rc = set_smem(pib_idx_kernel<1>, smem))) return rc;
        pib_idx_kernel<1><<<grid, PIB_THREADS, smem, st>>>(boxes, pts, out, T, M);
    }
    return check_launch("pib_idx_kernel");
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
