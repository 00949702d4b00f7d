// lg_strip.cuh -- the cull -> compact -> drain machinery shared by the IoU and NMS kernels.
//
// A CTA (256 threads; 512 for the lazy NMS) owns a block of the pair space.
//   cull    lanes along columns, row records broadcast from smem; (ca-cb)^2 > (ra+rb)^2 proves the
//           reference result is exactly +0.0.  The row iterations of a warp are fully unrolled, their
//           ballots kept in registers, so a warp does ONE shared atomicAdd per tile to reserve queue space;
//   queue   survivors of many tiles accumulate in one smem queue ((row, col) codes);
//   drain   every thread takes queue entries round-robin, so the expensive polygon path runs with full
//           warps whatever the survivor density (2-6 % for NMS, 0.3 % for anchors x GT, 100 % dense);
//   rare    pairs the fast path defers (> 8 polygon vertices or an angular near-tie, lg_geom.cuh) are
//           collected in a second queue and evaluated together by the literal path (drain_rare), as late
//           as the caller can afford, so that the 64 lanes that run it are well filled.
#pragma once
#include "lg_geom.cuh"

namespace lg {

constexpr int ST_THREADS = 256;
constexpr int ST_QCAP = 4096;                   // main queue entries
constexpr int ST_RARECAP = ST_QCAP + 2048;      // rare queue entries: a full main drain always fits on top of 2048
constexpr int ST_SLOW_LANES = 64;               // threads that run the literal path (each needs 4 slab columns)
constexpr size_t ST_SLAB_BYTES = (size_t)SLAB_ROWS_SMEM_B * ST_THREADS * sizeof(float2);  // 16 KB (column records in shared memory)
constexpr size_t ST_SLAB_BYTES_STAGED = (size_t)SLAB_ROWS * ST_THREADS * sizeof(float2);  // 24 KB: + 4 staged B corners per lane

// Reserve queue space for this warp's survivors of one tile and write their codes.
// m[k] = ballot of the k-th row iteration (row = rbase + rstep * k), code = row << SHIFT | col.
template <int SHIFT, int NK, typename Q>
__device__ __forceinline__ void push_survivors(const unsigned (&m)[NK], const int lane, const int rbase, const int rstep,
                                               const int col, int* qcount, Q* __restrict__ queue) {
    unsigned any = 0u;
#pragma unroll
    for (int k = 0; k < NK; k++) any |= m[k];
    if (any == 0u) return;
    int total = 0;
#pragma unroll
    for (int k = 0; k < NK; k++) total += __popc(m[k]);
    int base = 0;
    if (lane == 0) base = atomicAdd(qcount, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        if (m[k]) {
            if ((m[k] >> lane) & 1u) queue[base + __popc(m[k] & lt)] = (Q)(((rbase + rstep * k) << SHIFT) | col);
            base += __popc(m[k]);
        }
    }
}

// Drain `qn` queued pairs with the fast polygon path.  All threads of the CTA must call it; it contains no
// barrier, and the caller must place one before the queue (or *rcount) is touched again.
//   sA, sB   row / column records (shared memory, or global); code -> row = code >> SHIFT, col = code & (2^SHIFT - 1)
//   emit(row, col, overlap, A, B) consumes one result (store an IoU, set a mask bit, ...)
//   rareq    deferred pairs are appended at (*rcount)++; the caller guarantees *rcount + qn <= ST_RARECAP
template <int FL, int SHIFT, int NT = ST_THREADS, bool STAGE = false, typename Q, typename Emit>
__device__ __forceinline__ void drain_main(const float4* __restrict__ sA, const float4* __restrict__ sB, float2* __restrict__ slab,
                                           const Q* __restrict__ queue, const int qn, Q* __restrict__ rareq, int* rcount, Emit emit) {
    const int tid = threadIdx.x;
    for (int base = 0; base < qn; base += NT) {  // uniform trip count: the warp is converged at the ballot
        const int q = base + tid;
        const bool act = q < qn;
        const unsigned wm = __ballot_sync(0xffffffffu, act);
        if (act) {
            const unsigned e = queue[q];
            const int r = e >> SHIFT, c = e & ((1u << SHIFT) - 1u);
            const float4* A = sA + (size_t)r * REC_F4;
            const float4* B = sB + (size_t)c * REC_F4;
            const float ov = overlap_area<FL, STAGE>(A, B, slab + tid, NT, wm);
            if (ov < 0.f) rareq[atomicAdd(rcount, 1)] = (Q)e;
            else emit(r, c, ov, A, B);
        }
    }
}

// Evaluate the deferred pairs with the literal path.  The caller brackets it with barriers: one before (all
// pushes visible, the slab free) and one after (before *rcount, which is reset here by thread 0, is used again).
template <int FL, int SHIFT, int NT = ST_THREADS, typename Q, typename Emit>
__device__ __forceinline__ void drain_rare(const float4* __restrict__ sA, const float4* __restrict__ sB, float2* __restrict__ slab,
                                           const Q* __restrict__ rareq, int* rcount, Emit emit) {
    const int tid = threadIdx.x;
    const int rn = *rcount;
    if (rn > 0 && tid < ST_SLOW_LANES) {
        // vertex slots 0..7 -> column tid, 8..15 -> column tid + 64; angles -> columns tid + 128 (as floats)
        auto slab16 = [&](int k) -> float2& { return slab[(k & 7) * NT + tid + ((k >> 3) << 6)]; };
        auto ang16 = [&](int k) -> float& { return reinterpret_cast<float*>(slab + (k >> 1) * NT + 128 + tid)[k & 1]; };
        for (int q = tid; q < rn; q += ST_SLOW_LANES) {
            const unsigned e = rareq[q];
            const int r = e >> SHIFT, c = e & ((1u << SHIFT) - 1u);
            const float4* A = sA + (size_t)r * REC_F4;
            const float4* B = sB + (size_t)c * REC_F4;
            float ov = overlap_area16<FL>(A, B, slab16);
            if (ov < 0.f) ov = overlap_area_slow<FL>(A, B, slab16, ang16);  // angular near-tie: the reference's own procedure
            emit(r, c, ov, A, B);
        }
    }
    __syncthreads();
    if (tid == 0) *rcount = 0;
}

// ---- packed FP32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE single-precision operations per issue slot) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- warp-autonomous variant: every warp owns a private work list, no CTA barrier anywhere in the sweep ----
// Used by the N x M IoU strip kernel, whose warps have nothing to exchange: a warp appends the survivors of its
// own tile slice, runs the polygon path on full 32-entry rounds taken from the top of its list, and keeps the
// (< 32) rest for the next tile.  Deferred pairs go to a small per-warp list that 8 lanes work off.
constexpr int WQ_CAP = 31 + 512;  // leftover of the previous tiles + one tile slice (8 rows x 64 columns)
constexpr int WR_CAP = 64;

struct WarpQueue {
    uint32_t* list;  // WQ_CAP entries (shared memory)
    uint32_t* rare;  // WR_CAP entries
    int count, rcount;  // warp-uniform
};

template <int SHIFT, int NK>
__device__ __forceinline__ void warp_push(WarpQueue& q, const unsigned (&m)[NK], const int lane, const int rbase, const int rstep,
                                          const int col) {
    unsigned any = 0u;
#pragma unroll
    for (int k = 0; k < NK; k++) any |= m[k];
    if (any == 0u) return;
    const unsigned lt = (1u << lane) - 1u;
    int base = q.count;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        if (m[k]) {
            if ((m[k] >> lane) & 1u) q.list[base + __popc(m[k] & lt)] = (uint32_t)(((rbase + rstep * k) << SHIFT) | col);
            base += __popc(m[k]);
        }
    }
    q.count = base;
    __syncwarp();
}

// the deferred pairs of this warp: lanes 0..7, each with 16 vertex + 16 angle slots carved out of the warp's 32 slab columns
template <int FL, int SHIFT, typename Emit>
__device__ __forceinline__ void warp_drain_rare(WarpQueue& q, const float4* __restrict__ sA, const float4* __restrict__ sB,
                                                float2* __restrict__ slab_warp, const int sstride, const int lane, Emit emit) {
    __syncwarp();
    if (lane < 8) {
        auto slab16 = [&](int k) -> float2& { return slab_warp[(k & 7) * sstride + lane + ((k >> 3) << 3)]; };
        auto ang16 = [&](int k) -> float& { return reinterpret_cast<float*>(slab_warp + (k >> 1) * sstride + 16 + lane)[k & 1]; };
        for (int i = lane; i < q.rcount; i += 8) {
            const unsigned e = q.rare[i];
            const int r = e >> SHIFT, c = e & ((1u << SHIFT) - 1u);
            const float4* A = sA + (size_t)r * REC_F4;
            const float4* B = sB + (size_t)c * REC_F4;
            float ov = overlap_area16<FL>(A, B, slab16);
            if (ov < 0.f) ov = overlap_area_slow<FL>(A, B, slab16, ang16);
            emit(r, c, ov, A, B);
        }
    }
    q.rcount = 0;
    __syncwarp();
}

// run the polygon path on the top min(count, 32) entries of the warp's list
// INLINE_PATH: expand the polygon path in place (best when most pairs reach it) or call it out of line (keeps the
// caller's sweep loop in registers: best when most pairs are culled)
template <int FL, int SHIFT, bool INLINE_PATH, typename Emit>
__device__ __forceinline__ void warp_round(WarpQueue& q, const float4* __restrict__ sA, const float4* __restrict__ sB,
                                           float2* __restrict__ slab_warp, const int sstride, const int lane, Emit emit) {
    const int n = min(q.count, 32);
    const int idx = q.count - n + lane;
    const bool act = lane < n;
    const unsigned wm = __ballot_sync(0xffffffffu, act);
    bool defer = false;
    unsigned e = 0u;
    if (act) {
        e = q.list[idx];
        const int r = e >> SHIFT, c = e & ((1u << SHIFT) - 1u);
        const float4* A = sA + (size_t)r * REC_F4;
        const float4* B = sB + (size_t)c * REC_F4;
        const float ov = INLINE_PATH ? overlap_area<FL, true>(A, B, slab_warp + lane, sstride, wm) : overlap_area_call<FL>(A, B, slab_warp + lane, sstride, wm);
        if (ov < 0.f) defer = true;
        else emit(r, c, ov, A, B);
    }
    q.count -= n;
    const unsigned dm = __ballot_sync(0xffffffffu, defer);
    if (dm) {
        if (defer) q.rare[q.rcount + __popc(dm & ((1u << lane) - 1u))] = e;
        q.rcount += __popc(dm);
        if (q.rcount > WR_CAP - 32) warp_drain_rare<FL, SHIFT>(q, sA, sB, slab_warp, sstride, lane, emit);
    }
    __syncwarp();
}

}  // namespace lg
