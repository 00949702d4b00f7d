// lg_pool.cu -- the other users of check_pt_in_box3d, for sm_100a:
//   RoI-aware voxel pooling, forward and backward   (/root/reference/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:39-310)
//   RoI point pooling, forward                      (/root/reference/pcdet/ops/roipoint_pool3d/src/roipoint_pool3d_kernel.cu:15-164)
// The predicate (and the local coordinates it hands back) is lg_pib.cuh's, bit-for-bit the reference's.
//
// What is different from the reference's structure, and why the outputs are still identical:
//  * The reference materialises an (N boxes x M points) int mask in a cudaMalloc'ed scratch and then lets ONE THREAD PER
//    BOX walk all M mask entries to fill the voxel lists (collect_inside_pts_for_box3d) / the sample list
//    (get_pooled_idx).  Only the ORDER of that walk is semantics: a voxel keeps its first max_pts-1 points in ascending
//    point index, a box its first S.  Here one CTA owns a box: each warp tests a contiguous range of points and
//    ballot-compacts the inside ones (ascending within the warp), the per-warp lists are consumed in warp order
//    (ascending overall), and points that fall into the same voxel inside one 32-entry batch are ranked with
//    match.any -- the sequential insert, 32 at a time.  No mask, no scratch.
//  * Pooling: the reference runs a thread per (voxel, channel) with the channel in blockIdx.y, so neighbouring threads
//    write 4 bytes C*4 bytes apart.  The outputs are (voxel, channel) row-major: here every CTA first fills its slice
//    of pooled / argmax with the empty-voxel values in 16-byte stores (the outputs are HBM-write bound: >= 95 % of
//    the voxels of a 12^3 grid are empty) and then a lane group per NON-EMPTY voxel overwrites its C values.
//  * The library writes every output element (the reference needs its wrapper to zero-fill three tensors first).
#include "lg_common.cuh"
#include "lg_pib.cuh"

namespace lg {

constexpr int RA_THREADS = 512;
constexpr int RA_WARPS = RA_THREADS / 32;
constexpr int RA_WSEG = 1024;                  // points per warp per pass (10 bits of a packed list entry)
constexpr int RA_CHUNK = RA_WARPS * RA_WSEG;   // points per pass of the CTA
constexpr int RA_SMEM_VOX = 8192;              // voxel counters live in shared memory up to this many voxels
constexpr int RA_MAX_VOX = 1 << 22;            // 22 bits of a packed list entry

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// generate_pts_mask_for_box3d (kernel.cu:59-69): x_res = d / out; idx = int((local + d / 2) / x_res) converted to unsigned,
// then min(max(idx, 0), out - 1) on UNSIGNED operands: a negative conversion result ends up in the last voxel.
__device__ __forceinline__ unsigned voxel_axis(const float local, const float d, const int out) {
    const float res = __fdiv_rn(d, (float)out);
    const unsigned i = (unsigned)__float2int_rz(__fdiv_rn(__fadd_rn(local, __fmul_rn(d, 0.5f)), res));
    return min(i, (unsigned)(out - 1));
}

// the reference's check_pt_in_box3d incl. the local coordinates (lg_pib.cuh: pt_in_box)
template <int FL>
__device__ __forceinline__ bool pt_in_box_local(const float x, const float y, const float z, const float4 r0, const float4 r1, float& lx,
                                                float& ly) {
    if (fabsf(z - r0.z) > r0.w) return false;
    const float sx = x - r0.x, sy = y - r0.y;
    lx = msub<FL>(sx, r1.x, sy, r1.y);
    ly = madd_second<FL>(sx, r1.y, sy, r1.x);
    return (fabsf(lx) < r1.z) & (fabsf(ly) < r1.w);
}

// One CTA per box: voxel lists of pts_idx_of_voxels (N, V, max_pts), which the host has zero-filled.
template <int FL>
__global__ void __launch_bounds__(RA_THREADS, 1)
    roiaware_collect_kernel(const float* __restrict__ rois, const float* __restrict__ pts, const int m, const int ox, const int oy,
                            const int oz, const int max_pts, int32_t* __restrict__ pts_idx) {
    extern __shared__ uint32_t ra_sm[];
    uint32_t* wlist = ra_sm;                                   // [RA_WARPS][RA_WSEG] packed (voxel << 10 | offset in the warp's range)
    int* scnt = reinterpret_cast<int*>(ra_sm + RA_CHUNK);      // [V] when V <= RA_SMEM_VOX
    __shared__ int wcount[RA_WARPS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = ox * oy * oz;
    const bool smem_cnt = V <= RA_SMEM_VOX;
    const float* box = rois + (size_t)blockIdx.x * 7;
    int32_t* lists = pts_idx + (size_t)blockIdx.x * V * max_pts;
    float4 r0, r1;
    make_pib_record(box, 1e-5f, r0, r1);
    const float dx = box[3], dy = box[4], dz = box[5];
    const int cap = max_pts - 1;  // slot 0 is the counter
    if (smem_cnt)
        for (int v = tid; v < V; v += RA_THREADS) scnt[v] = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += RA_CHUNK) {
        // every warp: its contiguous range of points, inside ones compacted in ascending order
        const int64_t w0 = base + (int64_t)warp * RA_WSEG;
        int cnt = 0;
        for (int s = 0; s < RA_WSEG / 32 && w0 + s * 32 < m; ++s) {
            const int64_t k = w0 + s * 32 + lane;
            bool in = false;
            uint32_t v = 0;
            if (k < m) {
                const float x = __ldg(pts + k * 3), y = __ldg(pts + k * 3 + 1), z = __ldg(pts + k * 3 + 2);
                float lx, ly;
                in = pt_in_box_local<FL>(x, y, z, r0, r1, lx, ly);
                if (in) {
                    const float lz = z - r0.z;
                    v = (voxel_axis(lx, dx, ox) * oy + voxel_axis(ly, dy, oy)) * oz + voxel_axis(lz, dz, oz);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (in) wlist[warp * RA_WSEG + cnt + __popc(bal & lanemask_lt())] = (v << 10) | (uint32_t)(s * 32 + lane);
            cnt += __popc(bal);
        }
        if (lane == 0) wcount[warp] = cnt;
        __syncthreads();
        // warp 0: the reference's sequential insert (kernel.cu:88-101), 32 entries at a time, lists in warp order
        if (warp == 0) {
            for (int w = 0; w < RA_WARPS; ++w) {
                const int nw = wcount[w];
                for (int i0 = 0; i0 < nw; i0 += 32) {
                    const int i = i0 + lane;
                    const bool act = i < nw;
                    const uint32_t e = act ? wlist[w * RA_WSEG + i] : 0u;
                    const uint32_t v = e >> 10;
                    const unsigned peers = __match_any_sync(0xffffffffu, act ? v : 0xffffffe0u + lane);
                    const int rank = __popc(peers & lanemask_lt()), size = __popc(peers);
                    int* ctr = smem_cnt ? scnt + v : lists + (size_t)v * max_pts;
                    const int have = act ? *ctr : 0;
                    if (act && have + rank < cap) lists[(size_t)v * max_pts + have + rank + 1] = (int32_t)(base + w * RA_WSEG + (e & 1023u));
                    __syncwarp();
                    if (act && rank == size - 1) *ctr = min(have + size, cap);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
    }
    if (smem_cnt)
        for (int v = tid; v < V; v += RA_THREADS) {
            const int c = scnt[v];
            if (c) lists[(size_t)v * max_pts] = c;
        }
}

// Pooling over the voxel lists (roiaware_maxpool3d / roiaware_avgpool3d, kernel.cu:104-183).  METHOD 0 = max (+ argmax), 1 = avg.
// A CTA owns `vpc` consecutive voxels of the flattened (box, voxel) axis.
template <int METHOD>
__global__ void __launch_bounds__(256)
    roiaware_pool_kernel(const float* __restrict__ feat, const int32_t* __restrict__ pts_idx, const int64_t nv, const int c,
                         const int max_pts, const int gshift, const int vpc, float* __restrict__ pooled, int32_t* __restrict__ argmax) {
    const int tid = threadIdx.x;
    const int64_t v0 = (int64_t)blockIdx.x * vpc;
    const int64_t v1 = min(v0 + (int64_t)vpc, nv);
    // 1. empty-voxel values for the whole slice
    const int64_t e0 = v0 * c, ne = (v1 - v0) * c;
    const int fillv = METHOD == 0 ? -1 : 0;
    if ((c & 3) == 0) {
        float4* p4 = reinterpret_cast<float4*>(pooled + e0);
        int4* a4 = argmax ? reinterpret_cast<int4*>(argmax + e0) : nullptr;
        for (int64_t i = tid; i < (ne >> 2); i += 256) {
            __stcs(p4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
            if (a4) __stcs(a4 + i, make_int4(fillv, fillv, fillv, fillv));
        }
    } else {
        for (int64_t i = tid; i < ne; i += 256) {
            pooled[e0 + i] = 0.f;
            if (argmax) argmax[e0 + i] = fillv;
        }
    }
    __syncthreads();  // the overwrites below come after the fill, also for other threads' elements
    // 2. a group of 2^gshift lanes per non-empty voxel, channels strided over the group
    const int G = 1 << gshift, sub = tid & (G - 1);
    for (int64_t v = v0 + (tid >> gshift); v < v1; v += (256 >> gshift)) {
        const int32_t* l = pts_idx + v * max_pts;
        const int cnt = __ldg(l);
        if (cnt == 0) continue;
        for (int ch = sub; ch < c; ch += G) {
            if (METHOD == 0) {
                int am = -1;
                float mx = __int_as_float(0xff800000);  // float max_val = -1e50 (kernel.cu:131) is -inf
                for (int k = 1; k <= cnt; ++k) {
                    const int idx = __ldg(l + k);
                    const float f = __ldg(feat + (size_t)idx * c + ch);
                    if (f > mx) {
                        mx = f;
                        am = idx;
                    }
                }
                if (am != -1) pooled[v * c + ch] = mx;
                if (argmax) argmax[v * c + ch] = am;
            } else {
                float s = 0.f;
                for (int k = 1; k <= cnt; ++k) s = __fadd_rn(s, __ldg(feat + (size_t)__ldg(l + k) * c + ch));
                pooled[v * c + ch] = __fdiv_rn(s, (float)cnt);
            }
        }
    }
}

// roiaware_maxpool3d_backward / roiaware_avgpool3d_backward (kernel.cu:229-283): grad_in accumulates with float atomics (as the
// reference).  Empty voxels are skipped on their list counter alone -- one word per voxel instead of C argmax + C gradient
// values; an empty voxel has argmax -1 in every channel, so nothing is lost.
template <int METHOD>
__global__ void __launch_bounds__(256)
    roiaware_backward_kernel(const int32_t* __restrict__ pts_idx, const int32_t* __restrict__ argmax, const float* __restrict__ grad_out,
                             float* __restrict__ grad_in, const int64_t nv, const int c, const int max_pts, const int gshift) {
    const int G = 1 << gshift, sub = threadIdx.x & (G - 1);
    const int64_t stride = (int64_t)gridDim.x * (256 >> gshift);
    for (int64_t v = (int64_t)blockIdx.x * (256 >> gshift) + (threadIdx.x >> gshift); v < nv; v += stride) {
        const int32_t* l = pts_idx + v * max_pts;
        const int cnt = __ldg(l);
        if (cnt == 0) continue;
        for (int ch = sub; ch < c; ch += G) {
            const float g = __ldg(grad_out + v * c + ch);
            if (METHOD == 0) {
                const int a = __ldg(argmax + v * c + ch);
                if (a != -1) atomicAdd(grad_in + (size_t)a * c + ch, g);
            } else {
                const float val = __fmul_rn(g, __fdiv_rn(1.f, fmaxf((float)cnt, 1.f)));
                for (int k = 1; k <= cnt; ++k) atomicAdd(grad_in + (size_t)__ldg(l + k) * c + ch, val);
            }
        }
    }
}

// ---- RoI point pooling ----------------------------------------------------------------------------
constexpr int RP_THREADS = 512;
constexpr int RP_WARPS = RP_THREADS / 32;

// One CTA per (box, frame): the first S inside points in ascending index (get_pooled_idx, kernel.cu:63-98), repeated cyclically
// when there are fewer, then one coalesced gather of (xyz, features) rows (roipool3d_forward, kernel.cu:101-134).
template <int FL>
__global__ void __launch_bounds__(RP_THREADS, 1)
    roipoint_pool_kernel(const float* __restrict__ xyz, const float* __restrict__ boxes, const float* __restrict__ feat, const int n,
                         const int m, const int c, const int S, float* __restrict__ pooled, int32_t* __restrict__ empty_flag) {
    extern __shared__ int32_t rp_sm[];
    int32_t* wlist = rp_sm;                  // [RP_WARPS][S]
    int32_t* sidx = rp_sm + RP_WARPS * S;    // [S]
    __shared__ int wstart[RP_WARPS + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t bj = (size_t)blockIdx.y * m + blockIdx.x;
    const float* P = xyz + (size_t)blockIdx.y * n * 3;
    const float* F = feat + (size_t)blockIdx.y * n * c;
    float4 r0, r1;
    make_pib_record(boxes + bj * 7, 1e-5f, r0, r1);
    const int seg = ((n + RP_WARPS - 1) / RP_WARPS + 31) & ~31;  // points per warp, whole warp steps
    const int k0 = warp * seg, k1 = min(k0 + seg, n);
    int cnt = 0;
    for (int kb = k0; kb < k1 && cnt < S; kb += 32) {
        const int k = kb + lane;
        bool in = false;
        if (k < k1) {
            float lx, ly;
            in = pt_in_box_local<FL>(__ldg(P + (size_t)k * 3), __ldg(P + (size_t)k * 3 + 1), __ldg(P + (size_t)k * 3 + 2), r0, r1, lx, ly);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int pos = cnt + __popc(bal & lanemask_lt());
        if (in && pos < S) wlist[warp * S + pos] = k;
        cnt = min(cnt + __popc(bal), S);
    }
    if (lane == 0) wstart[warp + 1] = cnt;
    __syncthreads();
    if (tid == 0) {
        wstart[0] = 0;
        for (int w = 0; w < RP_WARPS; ++w) wstart[w + 1] += wstart[w];
    }
    __syncthreads();
    const int total = min(wstart[RP_WARPS], S);
    float* dst = pooled + bj * (size_t)S * (3 + c);
    const int W = 3 + c;
    const int64_t ne = (int64_t)S * W;
    if (total == 0) {  // the reference leaves the wrapper's zeros in place and raises the flag
        if (tid == 0) empty_flag[bj] = 1;
        for (int64_t e = tid; e < ne; e += RP_THREADS) dst[e] = 0.f;
        return;
    }
    if (tid == 0) empty_flag[bj] = 0;
    for (int t = tid; t < S; t += RP_THREADS) {
        const int p = t < total ? t : t % total;  // kernel.cu:90-96
        int w = 0;
        while (p >= wstart[w + 1]) ++w;
        sidx[t] = wlist[w * S + (p - wstart[w])];
    }
    __syncthreads();
    for (int64_t e = tid; e < ne; e += RP_THREADS) {
        const int row = (int)(e / W), col = (int)(e - (int64_t)row * W);
        const size_t src = (size_t)sidx[row];
        dst[e] = col < 3 ? __ldg(P + src * 3 + col) : __ldg(F + src * c + (col - 3));
    }
}

static int group_shift(int c) {  // lanes per voxel: the power of two >= C, at most a warp
    int s = 0;
    while ((1 << s) < c && s < 5) ++s;
    return s;
}

}  // namespace lg

extern "C" int lg_roiaware_pool3d_forward(const float* rois, int num_rois, const float* pts, int num_pts, const float* pts_feature,
                                          int channels, int out_x, int out_y, int out_z, int max_pts_each_voxel, int pool_method,
                                          float* pooled_features, int32_t* argmax, int32_t* pts_idx_of_voxels, unsigned flags,
                                          void* stream) {
    using namespace lg;
    if (num_rois < 0 || num_pts < 0 || channels < 0 || out_x < 1 || out_y < 1 || out_z < 1 || out_x > 255 || out_y > 255 || out_z > 255 ||
        max_pts_each_voxel < 1 || (pool_method != 0 && pool_method != 1)) {
        set_error("roiaware_pool3d: bad argument (rois=%d pts=%d channels=%d out=%dx%dx%d max_pts=%d method=%d; out sizes are 1..255)",
                  num_rois, num_pts, channels, out_x, out_y, out_z, max_pts_each_voxel, pool_method);
        return LG_ERR_INVALID_ARG;
    }
    if (num_rois == 0) return LG_OK;
    const int64_t V = (int64_t)out_x * out_y * out_z;
    if (V > RA_MAX_VOX) {
        set_error("roiaware_pool3d: %lld voxels per box exceed the limit of %d", (long long)V, RA_MAX_VOX);
        return LG_ERR_TOO_LARGE;
    }
    if (!rois || !pts_idx_of_voxels || (num_pts > 0 && !pts) || (channels > 0 && (!pooled_features || (num_pts > 0 && !pts_feature)))) {
        set_error("roiaware_pool3d: null pointer");
        return LG_ERR_INVALID_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t nv = (int64_t)num_rois * V;
    cudaError_t e = cudaMemsetAsync(pts_idx_of_voxels, 0, (size_t)nv * max_pts_each_voxel * sizeof(int32_t), st);
    if (e != cudaSuccess) {
        set_error("roiaware_pool3d: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return (int)e;
    }
    int rc;
    if (num_pts > 0) {
        const size_t smem = (size_t)RA_CHUNK * 4 + (V <= RA_SMEM_VOX ? (size_t)V * 4 : 0);
        if (flags & LG_FLAG_STRICT_FP32) {
            if ((rc = set_smem(roiaware_collect_kernel<0>, smem))) return rc;
            roiaware_collect_kernel<0><<<num_rois, RA_THREADS, smem, st>>>(rois, pts, num_pts, out_x, out_y, out_z, max_pts_each_voxel,
                                                                          pts_idx_of_voxels);
        } else {
            if ((rc = set_smem(roiaware_collect_kernel<1>, smem))) return rc;
            roiaware_collect_kernel<1><<<num_rois, RA_THREADS, smem, st>>>(rois, pts, num_pts, out_x, out_y, out_z, max_pts_each_voxel,
                                                                          pts_idx_of_voxels);
        }
        if ((rc = check_launch("roiaware_collect_kernel"))) return rc;
    }
    if (channels == 0) return LG_OK;
    int vpc = 16384 / channels;  // ~64 KB of pooled values per CTA
    vpc = vpc < 32 ? 32 : (vpc > 4096 ? 4096 : vpc);
    const int64_t grid = (nv + vpc - 1) / vpc;
    if (grid > 0x7fffffffLL) {
        set_error("roiaware_pool3d: too many voxels for one launch");
        return LG_ERR_TOO_LARGE;
    }
    const int gs = group_shift(channels);
    if (pool_method == 0)
        roiaware_pool_kernel<0><<<(unsigned)grid, 256, 0, st>>>(pts_feature, pts_idx_of_voxels, nv, channels, max_pts_each_voxel, gs, vpc,
                                                               pooled_features, argmax);
    else
        roiaware_pool_kernel<1><<<(unsigned)grid, 256, 0, st>>>(pts_feature, pts_idx_of_voxels, nv, channels, max_pts_each_voxel, gs, vpc,
                                                               pooled_features, argmax);
    return check_launch("roiaware_pool_kernel");
}

extern "C" int lg_roiaware_pool3d_backward(const int32_t* pts_idx_of_voxels, const int32_t* argmax, const float* grad_out, float* grad_in,
                                           int num_rois, int out_x, int out_y, int out_z, int channels, int max_pts_each_voxel,
                                           int pool_method, unsigned, void* stream) {
    using namespace lg;
    if (num_rois < 0 || channels < 0 || out_x < 1 || out_y < 1 || out_z < 1 || max_pts_each_voxel < 1 || (pool_method != 0 && pool_method != 1)) {
        set_error("roiaware_pool3d_backward: bad argument");
        return LG_ERR_INVALID_ARG;
    }
    if (num_rois == 0 || channels == 0) return LG_OK;
    if (!pts_idx_of_voxels || !grad_out || !grad_in || (pool_method == 0 && !argmax)) {
        set_error("roiaware_pool3d_backward: null pointer");
        return LG_ERR_INVALID_ARG;
    }
    const int64_t nv = (int64_t)num_rois * out_x * out_y * out_z;
    const int gs = group_shift(channels);
    const int64_t per_cta = 256 >> gs;
    int64_t grid = (nv + per_cta - 1) / per_cta;
    if (grid > 148 * 64) grid = 148 * 64;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (pool_method == 0)
        roiaware_backward_kernel<0><<<(unsigned)grid, 256, 0, st>>>(pts_idx_of_voxels, argmax, grad_out, grad_in, nv, channels, max_pts_each_voxel, gs);
    else
        roiaware_backward_kernel<1><<<(unsigned)grid, 256, 0, st>>>(pts_idx_of_voxels, argmax, grad_out, grad_in, nv, channels, max_pts_each_voxel, gs);
    return check_launch("roiaware_backward_kernel");
}

extern "C" int lg_roipoint_pool3d_forward(const float* xyz, const float* boxes3d, const float* pts_feature, int batch, int num_pts,
                                          int num_boxes, int channels, int num_sampled, float* pooled_features, int32_t* pooled_empty_flag,
                                          unsigned flags, void* stream) {
    using namespace lg;
    if (batch < 0 || num_pts < 0 || num_boxes < 0 || channels < 0 || num_sampled < 0) {
        set_error("roipoint_pool3d: negative size");
        return LG_ERR_INVALID_ARG;
    }
    if (batch == 0 || num_boxes == 0) return LG_OK;
    if (num_sampled > LG_ROIPOINT_MAX_SAMPLES) {
        set_error("roipoint_pool3d: num_sampled_points=%d exceeds LG_ROIPOINT_MAX_SAMPLES=%d", num_sampled, LG_ROIPOINT_MAX_SAMPLES);
        return LG_ERR_TOO_LARGE;
    }
    if (batch > 65535) {
        set_error("roipoint_pool3d: batch=%d exceeds 65535; split the batch", batch);
        return LG_ERR_TOO_LARGE;
    }
    if (!boxes3d || !pooled_empty_flag || (num_sampled > 0 && !pooled_features) || (num_pts > 0 && (!xyz || (channels > 0 && !pts_feature)))) {
        set_error("roipoint_pool3d: null pointer");
        return LG_ERR_INVALID_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)(RP_WARPS + 1) * (num_sampled > 0 ? num_sampled : 1) * sizeof(int32_t);
    dim3 grid((unsigned)num_boxes, (unsigned)batch);
    int rc;
    if (flags & LG_FLAG_STRICT_FP32) {
        if ((rc = set_smem(roipoint_pool_kernel<0>, smem))) return rc;
        roipoint_pool_kernel<0><<<grid, RP_THREADS, smem, st>>>(xyz, boxes3d, pts_feature, num_pts, num_boxes, channels, num_sampled,
                                                               pooled_features, pooled_empty_flag);
    } else {
        if ((rc = set_smem(roipoint_pool_kernel<1>, smem))) return rc;
        roipoint_pool_kernel<1><<<grid, RP_THREADS, smem, st>>>(xyz, boxes3d, pts_feature, num_pts, num_boxes, channels, num_sampled,
                                                               pooled_features, pooled_empty_flag);
    }
    return check_launch("roipoint_pool_kernel");
}
