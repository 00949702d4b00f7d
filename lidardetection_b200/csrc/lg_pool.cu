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
//    write 4 bytes C*4 bytes apart, and it visits every voxel.  >= 95 % of the voxels of a 12^3 grid are empty, so the
//    three outputs are HBM-write bound constants: one fill kernel writes the empty-voxel values (lists 0, pooled 0,
//    argmax -1) over the whole machine in 16-byte stores, and the box's CTA, which knows its non-empty voxels from the
//    counters it keeps in shared memory, pools only those right after the insert (a lane group per voxel, coalesced
//    along the channels).  The library writes every output element (the reference needs its wrapper to zero-fill them).
//  * RoI point pooling: a box's S output rows are gathered by a thread-block cluster of up to 8 CTAs; each CTA tests
//    1/8 of the points, the per-warp lists are read across the cluster through distributed shared memory.
#include <cooperative_groups.h>

#include "lg_common.cuh"
#include "lg_pib.cuh"

namespace cg = cooperative_groups;

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

// Empty-voxel values for the three outputs of the forward, over the whole machine.
__device__ __forceinline__ void fill_words(int32_t* __restrict__ p, const size_t n, const int v) {
    if (!p || n == 0) return;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    const size_t head = min(n, (size_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / 4);  // words up to 16-byte alignment
    int4* p4 = reinterpret_cast<int4*>(p + head);
    const size_t n4 = (n - head) / 4;
    for (size_t i = gtid; i < n4; i += gsz) __stcs(p4 + i, make_int4(v, v, v, v));
    if (gtid < head) p[gtid] = v;
    const size_t tail0 = head + n4 * 4;
    if (gtid < n - tail0) p[tail0 + gtid] = v;
}

__global__ void __launch_bounds__(256)
    roiaware_fill_kernel(int32_t* __restrict__ lists, const size_t n_lists, int32_t* __restrict__ pooled, const size_t n_pooled,
                         int32_t* __restrict__ argmax, const int argmax_fill) {
    fill_words(lists, n_lists, 0);
    fill_words(pooled, n_pooled, 0);  // +0.0f
    fill_words(argmax, n_pooled, argmax_fill);
}

// Pooling of one box's non-empty voxels by its CTA.  A group of 2^gshift lanes owns a voxel; a lane owns up to four units of
// VEC consecutive channels (unit u0 + j * G, so that the group's loads of one j are contiguous) and walks the voxel's points in
// list order with the four loads of a point in flight together -- the loop is bound by the latency of the dependent
// list -> feature-row loads, so the group is kept as narrow as four units per lane allow (8 lanes for 128 channels) and many
// voxels are in flight per warp.  max: strict `>` in list order (first maximum wins; NaN and -inf never win, kernel.cu:131-139);
// avg: sum in list order, one IEEE division (kernel.cu:172-182).
template <int METHOD, int VEC>
__device__ __forceinline__ void pool_voxels(const int32_t* lists, const int* scnt, const float* __restrict__ feat, const int V, const int c,
                                            const int max_pts, const int gshift, const size_t vbase, float* __restrict__ pooled,
                                            int32_t* __restrict__ argmax) {
    const int tid = threadIdx.x, G = 1 << gshift, sub = tid & (G - 1);
    const int nunits = c / VEC;
    for (int v = tid >> gshift; v < V; v += (RA_THREADS >> gshift)) {
        const int32_t* l = lists + (size_t)v * max_pts;
        const int cnt = scnt ? scnt[v] : *l;
        if (cnt == 0) continue;
        if (scnt && sub == 0) *const_cast<int32_t*>(l) = cnt;
        for (int u0 = sub; u0 < nunits; u0 += 4 * G) {
            float acc[4][VEC];
            int am[4][VEC];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int t = 0; t < VEC; ++t) {
                    acc[j][t] = METHOD == 0 ? __int_as_float(0xff800000) : 0.f;  // float max_val = -1e50 is -inf
                    am[j][t] = -1;
                }
            for (int k = 1; k <= cnt; ++k) {
                const int idx = l[k];
                const float* row = feat + (size_t)idx * c;
                float f[4][VEC];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = u0 + j * G;
                    if (u < nunits) {
                        if (VEC == 4) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(row) + u);
                            f[j][0] = q.x, f[j][1 % VEC] = q.y, f[j][2 % VEC] = q.z, f[j][3 % VEC] = q.w;
                        } else {
                            f[j][0] = __ldg(row + u);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (u0 + j * G < nunits) {
#pragma unroll
                        for (int t = 0; t < VEC; ++t) {
                            if (METHOD == 0) {
                                if (f[j][t] > acc[j][t]) {
                                    acc[j][t] = f[j][t];
                                    am[j][t] = idx;
                                }
                            } else {
                                acc[j][t] = __fadd_rn(acc[j][t], f[j][t]);
                            }
                        }
                    }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = u0 + j * G;
                if (u >= nunits) continue;
                const size_t o = (vbase + v) * c + (size_t)u * VEC;
                if (METHOD == 0) {
#pragma unroll
                    for (int t = 0; t < VEC; ++t)
                        if (am[j][t] == -1) acc[j][t] = 0.f;  // the reference leaves the wrapper's zero in place
                    if (VEC == 4) {
                        *reinterpret_cast<float4*>(pooled + o) = make_float4(acc[j][0], acc[j][1 % VEC], acc[j][2 % VEC], acc[j][3 % VEC]);
                        *reinterpret_cast<int4*>(argmax + o) = make_int4(am[j][0], am[j][1 % VEC], am[j][2 % VEC], am[j][3 % VEC]);
                    } else {
                        pooled[o] = acc[j][0];
                        argmax[o] = am[j][0];
                    }
                } else {
                    const float n = (float)cnt;
                    if (VEC == 4)
                        *reinterpret_cast<float4*>(pooled + o) = make_float4(__fdiv_rn(acc[j][0], n), __fdiv_rn(acc[j][1 % VEC], n),
                                                                             __fdiv_rn(acc[j][2 % VEC], n), __fdiv_rn(acc[j][3 % VEC], n));
                    else
                        pooled[o] = __fdiv_rn(acc[j][0], n);
                }
            }
        }
    }
}

// One CTA per box: voxel lists into pts_idx_of_voxels (N, V, max_pts) (pre-filled with zeros), then pooling of the box's
// non-empty voxels into pooled / argmax (pre-filled with the empty-voxel values).  METHOD 0 = max (+ argmax), 1 = avg
// (roiaware_maxpool3d / roiaware_avgpool3d, kernel.cu:104-183).
template <int FL, int METHOD>
__global__ void __launch_bounds__(RA_THREADS, 1)
    roiaware_collect_pool_kernel(const float* __restrict__ rois, const float* __restrict__ pts, const float* __restrict__ feat,
                                 const int m, const int c, const int ox, const int oy, const int oz, const int max_pts, const int gshift,
                                 int32_t* __restrict__ pts_idx, float* __restrict__ pooled, int32_t* __restrict__ argmax) {
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
    make_pib_record<1>(box, 1e-5f, r0, r1);
    const float dx = box[3], dy = box[4], dz = box[5];
    const int cap = max_pts - 1;  // slot 0 is the counter
    if (smem_cnt)
        for (int v = tid; v < V; v += RA_THREADS) scnt[v] = 0;
    __syncthreads();
    for (int64_t base = 0; base < m; base += RA_CHUNK) {
        // every warp: its contiguous range of points, inside ones compacted in ascending order; four warp steps per
        // iteration with all twelve loads issued first (the loop is latency bound: the points come from L2)
        const int64_t w0 = base + (int64_t)warp * RA_WSEG;
        int cnt = 0;
        for (int s = 0; s < RA_WSEG / 32 && w0 + s * 32 < m; s += 4) {
            float x[4], y[4], z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t k = w0 + (s + u) * 32 + lane;
                const bool ok = k < m;
                x[u] = ok ? __ldg(pts + k * 3) : 0.f;
                y[u] = ok ? __ldg(pts + k * 3 + 1) : 0.f;
                z[u] = ok ? __ldg(pts + k * 3 + 2) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t k = w0 + (s + u) * 32 + lane;
                bool in = false;
                uint32_t v = 0;
                if (k < m) {
                    float lx, ly;
                    in = pt_in_box_local<FL>(x[u], y[u], z[u], r0, r1, lx, ly);
                    if (in) {
                        const float lz = z[u] - r0.z;
                        v = (voxel_axis(lx, dx, ox) * oy + voxel_axis(ly, dy, oy)) * oz + voxel_axis(lz, dz, oz);
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) wlist[warp * RA_WSEG + cnt + __popc(bal & lanemask_lt())] = (v << 10) | (uint32_t)((s + u) * 32 + lane);
                cnt += __popc(bal);
            }
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
        __syncthreads();  // also makes the list entries visible to the pooling below
    }
    // counters out, and pooling of the non-empty voxels
    if ((c & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(pooled) & 15) == 0 &&
        (METHOD == 1 || (reinterpret_cast<uintptr_t>(argmax) & 15) == 0))
        pool_voxels<METHOD, 4>(lists, smem_cnt ? scnt : nullptr, feat, V, c, max_pts, gshift, (size_t)blockIdx.x * V, pooled, argmax);
    else
        pool_voxels<METHOD, 1>(lists, smem_cnt ? scnt : nullptr, feat, V, c, max_pts, gshift, (size_t)blockIdx.x * V, pooled, argmax);
}

// roiaware_maxpool3d_backward / roiaware_avgpool3d_backward (kernel.cu:229-283): grad_in accumulates with float atomics (as the
// reference).  A thread reads the list counter of one voxel -- one word per voxel instead of C argmax + C gradient values; an
// empty voxel has argmax -1 in every channel, so nothing is lost -- and the warp then works off its non-empty voxels together.
template <int METHOD>
__global__ void __launch_bounds__(256)
    roiaware_backward_kernel(const int32_t* __restrict__ pts_idx, const int32_t* __restrict__ argmax, const float* __restrict__ grad_out,
                             float* __restrict__ grad_in, const int64_t nv, const int c, const int max_pts) {
    const int lane = threadIdx.x & 31;
    const int64_t v_own = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int cnt_own = v_own < nv ? __ldg(pts_idx + v_own * max_pts) : 0;
    unsigned todo = __ballot_sync(0xffffffffu, cnt_own != 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t v = v_own - lane + src;
        const int cnt = __shfl_sync(0xffffffffu, cnt_own, src);
        const int32_t* l = pts_idx + v * max_pts;
        for (int ch = lane; ch < c; ch += 32) {
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
constexpr int RP_MAX_CLUSTER = 8;

// A cluster of gridDim.z CTAs per (box, frame): the first S inside points in ascending index (get_pooled_idx, kernel.cu:63-98),
// repeated cyclically when there are fewer, then a coalesced gather of (xyz, features) rows (roipool3d_forward, kernel.cu:101-134).
// CTA r of the cluster tests the r-th contiguous slice of the points (each warp a contiguous range, so cluster rank major /
// warp minor is ascending point order) and gathers the r-th slice of the S output rows; lists and counts are read across the
// cluster through distributed shared memory.
template <int FL>
__global__ void __launch_bounds__(RP_THREADS, 2)
    roipoint_pool_kernel(const float* __restrict__ xyz, const float* __restrict__ boxes, const float* __restrict__ feat, const int n,
                         const int m, const int c, const int S, const int seg, const int wcap, float* __restrict__ pooled,
                         int32_t* __restrict__ empty_flag) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)gridDim.z, crank = (int)blockIdx.z;
    extern __shared__ int32_t rp_sm[];
    int32_t* wlist = rp_sm;                     // [RP_WARPS][wcap]
    int32_t* sidx = rp_sm + RP_WARPS * wcap;    // [rows of this CTA]
    __shared__ int wcnt[RP_WARPS];
    __shared__ int pre[RP_MAX_CLUSTER * RP_WARPS + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t bj = (size_t)blockIdx.y * m + blockIdx.x;
    const float* P = xyz + (size_t)blockIdx.y * n * 3;
    const float* F = feat + (size_t)blockIdx.y * n * c;
    float4 r0, r1;
    make_pib_record<1>(boxes + bj * 7, 1e-5f, r0, r1);
    const int64_t k0 = ((int64_t)crank * RP_WARPS + warp) * seg;
    const int k1 = (int)min(k0 + seg, (int64_t)n);
    int cnt = 0;
    for (int kb = (int)min(k0, (int64_t)n); kb < k1 && cnt < wcap; kb += 128) {
        float x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = kb + u * 32 + lane;
            const bool ok = k < k1;
            x[u] = ok ? __ldg(P + (size_t)k * 3) : 0.f;
            y[u] = ok ? __ldg(P + (size_t)k * 3 + 1) : 0.f;
            z[u] = ok ? __ldg(P + (size_t)k * 3 + 2) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = kb + u * 32 + lane;
            float lx, ly;
            const bool in = k < k1 && pt_in_box_local<FL>(x[u], y[u], z[u], r0, r1, lx, ly);
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            const int pos = cnt + __popc(bal & lanemask_lt());
            if (in && pos < wcap) wlist[warp * wcap + pos] = k;
            cnt = min(cnt + __popc(bal), wcap);
        }
    }
    if (lane == 0) wcnt[warp] = cnt;
    cluster.sync();
    if (tid < C * RP_WARPS) pre[tid + 1] = *cluster.map_shared_rank(&wcnt[tid & (RP_WARPS - 1)], tid / RP_WARPS);
    __syncthreads();
    if (tid == 0) {
        pre[0] = 0;
        for (int j = 0; j < C * RP_WARPS; ++j) pre[j + 1] += pre[j];
    }
    __syncthreads();
    const int total = min(pre[C * RP_WARPS], S);
    const int rows_per = (S + C - 1) / C;
    const int row0 = min(crank * rows_per, S), row1 = min(row0 + rows_per, S);
    const int W = 3 + c;
    float* dst = pooled + bj * (size_t)S * W;
    if (tid == 0 && crank == 0) empty_flag[bj] = total == 0;
    if (total > 0)
        for (int t = row0 + tid; t < row1; t += RP_THREADS) {
            const int p = t < total ? t : t % total;  // kernel.cu:90-96
            int j = 0;
            while (p >= pre[j + 1]) ++j;
            sidx[t - row0] = *cluster.map_shared_rank(&wlist[(j & (RP_WARPS - 1)) * wcap + (p - pre[j])], j / RP_WARPS);
        }
    cluster.sync();  // nobody leaves (or overwrites anything) while its lists may still be read
    // gather: a warp per output row, lanes along the row (stores of consecutive rows are contiguous)
    for (int row = row0 + warp; row < row1; row += RP_WARPS) {
        float* d = dst + (size_t)row * W;
        if (total == 0) {  // the reference leaves the wrapper's zeros in place
            for (int col = lane; col < W; col += 32) d[col] = 0.f;
            continue;
        }
        const size_t src = (size_t)sidx[row - row0];
        for (int col = lane; col < W; col += 32) __stcs(d + col, col < 3 ? __ldg(P + src * 3 + col) : __ldg(F + src * c + (col - 3)));
    }
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
    if (!rois || !pts_idx_of_voxels || (num_pts > 0 && !pts) ||
        (channels > 0 && (!pooled_features || (pool_method == 0 && !argmax) || (num_pts > 0 && !pts_feature)))) {
        set_error("roiaware_pool3d: null pointer");
        return LG_ERR_INVALID_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t nv = (int64_t)num_rois * V;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    roiaware_fill_kernel<<<sms * 8, 256, 0, st>>>(pts_idx_of_voxels, (size_t)nv * max_pts_each_voxel, reinterpret_cast<int32_t*>(pooled_features),
                                                  (size_t)nv * channels, channels > 0 ? argmax : nullptr, pool_method == 0 ? -1 : 0);
    int rc;
    if ((rc = check_launch("roiaware_fill_kernel"))) return rc;
    const size_t smem = (size_t)RA_CHUNK * 4 + (V <= RA_SMEM_VOX ? (size_t)V * 4 : 0);
    const int units = (channels & 3) == 0 ? channels / 4 : channels;  // float4 units when the rows allow it (pool_voxels)
    int gs = 0;  // lanes per voxel: four units per lane, at most a warp
    while ((1 << gs) * 4 < units && gs < 5) ++gs;
    const bool strict = flags & LG_FLAG_STRICT_FP32;
    auto launch = [&](auto kern) -> int {
        int r = set_smem(kern, smem);
        if (r) return r;
        kern<<<num_rois, RA_THREADS, smem, st>>>(rois, pts, pts_feature, num_pts, channels, out_x, out_y, out_z, max_pts_each_voxel, gs,
                                                pts_idx_of_voxels, pooled_features, argmax);
        return check_launch("roiaware_collect_pool_kernel");
    };
    if (pool_method == 0) return strict ? launch(roiaware_collect_pool_kernel<0, 0>) : launch(roiaware_collect_pool_kernel<1, 0>);
    return strict ? launch(roiaware_collect_pool_kernel<0, 1>) : launch(roiaware_collect_pool_kernel<1, 1>);
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
    const int64_t grid = (nv + 255) / 256;
    if (grid > 0x7fffffffLL) {
        set_error("roiaware_pool3d_backward: too many voxels for one launch");
        return LG_ERR_TOO_LARGE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (pool_method == 0)
        roiaware_backward_kernel<0><<<(unsigned)grid, 256, 0, st>>>(pts_idx_of_voxels, argmax, grad_out, grad_in, nv, channels, max_pts_each_voxel);
    else
        roiaware_backward_kernel<1><<<(unsigned)grid, 256, 0, st>>>(pts_idx_of_voxels, argmax, grad_out, grad_in, nv, channels, max_pts_each_voxel);
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
    // CTAs per box (one cluster): enough that the gather keeps ~8 CTAs per SM busy, at least 16 rows and 512 points each
    int split = 1;
    while (split < RP_MAX_CLUSTER && (int64_t)num_boxes * batch * split < 148 * 8 && num_sampled / (split * 2) >= 16 && num_pts / (split * 2) >= 512)
        split *= 2;
    const int64_t per_warp = ((int64_t)num_pts + (int64_t)split * RP_WARPS - 1) / ((int64_t)split * RP_WARPS);
    const int seg = (int)((per_warp + 127) / 128 * 128);  // points per warp: whole 4-step iterations
    const int wcap = num_sampled < seg ? num_sampled : seg;
    const int rows_per = (num_sampled + split - 1) / split;
    const size_t smem = ((size_t)RP_WARPS * wcap + rows_per + 1) * sizeof(int32_t);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)num_boxes, (unsigned)batch, (unsigned)split);
    lc.blockDim = dim3(RP_THREADS);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = (unsigned)split;
    lc.attrs = at;
    lc.numAttrs = 1;
    auto launch = [&](auto kern) -> int {
        int r = set_smem(kern, smem);
        if (r) return r;
        cudaError_t e = cudaLaunchKernelEx(&lc, kern, xyz, boxes3d, pts_feature, num_pts, num_boxes, channels, num_sampled, seg, wcap,
                                           pooled_features, pooled_empty_flag);
        if (e != cudaSuccess) {
            set_error("roipoint_pool_kernel: %s", cudaGetErrorString(e));
            return (int)e;
        }
        return check_launch("roipoint_pool_kernel");
    };
    return (flags & LG_FLAG_STRICT_FP32) ? launch(roipoint_pool_kernel<0>) : launch(roipoint_pool_kernel<1>);
}
