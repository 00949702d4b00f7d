// lg_strip.cuh -- the cull -> compact -> drain machinery shared by the IoU and NMS kernels.
//
// A CTA (256 threads) owns a block of the pair space whose box records sit in shared memory.
//   cull    lanes along columns, row records broadcast from smem; (ca-cb)^2 > (ra+rb)^2 proves the
//           reference result is exactly +0.0.  The row iterations of a warp are fully unrolled, their
//           ballots kept in registers, so a warp does ONE shared atomicAdd per tile to reserve queue space;
//   queue   survivors of several tiles accumulate in one smem queue (16-bit (row, col) codes);
//   drain   every thread takes queue entries round-robin, so the expensive polygon path runs with full
//           warps whatever the survivor density (2-6 % for NMS, 0.3 % for anchors x GT, 100 % dense);
//   rare    pairs the fast path defers (> 8 polygon vertices or an angular near-tie, lg_geom.cuh) are
//           collected in a second queue and evaluated together at the end of the drain by the literal path.
#pragma once
#include "lg_geom.cuh"

namespace lg {

constexpr int ST_THREADS = 256;
constexpr int ST_QCAP = 4096;      // queue entries (u16 codes); also the capacity of the rare queue
constexpr int ST_SLOW_LANES = 64;  // threads that run the literal path (each needs 4 slab columns)

struct DrainSmem {
    static constexpr size_t slab_bytes = (size_t)8 * ST_THREADS * sizeof(float2);  // 16 KB
    static constexpr size_t queue_bytes = (size_t)ST_QCAP * sizeof(uint16_t);      // 8 KB
    static constexpr size_t total = slab_bytes + 2 * queue_bytes;
};

// Reserve queue space for this warp's survivors of one tile and write their codes.
// m[k] = ballot of the k-th row iteration (row = rbase + rstep * k), code = row << SHIFT | col.
template <int SHIFT, int NK, typename Q>
__device__ __forceinline__ void push_survivors(const unsigned (&m)[NK], const int lane, const int rbase, const int rstep,
                                               const int col, int* qcount, Q* __restrict__ queue) {
    int total = 0;
#pragma unroll
    for (int k = 0; k < NK; k++) total += __popc(m[k]);
    if (total == 0) return;
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

// Drain `qn` queued pairs.  Must be called by all threads of the CTA (it contains __syncthreads()).
//   sA, sB   row / column records (shared memory, or global for the lazy NMS columns); a code decodes to row = code >> SHIFT, col = code & (2^SHIFT - 1)
//   emit(row, col, overlap, A, B) consumes one result (store an IoU, set a mask bit, ...)
//   *rcount  must be 0 on entry (made visible by a barrier); it is reset to 0 on exit, and the caller must
//            pass a barrier before the next drain.
//   Q        code type: uint16_t (row, col inside a tile) or uint32_t (lazy NMS: candidate slot, box index)
template <int FL, int SHIFT, typename Q, typename Emit>
__device__ __forceinline__ void drain_pairs(const float4* __restrict__ sA, const float4* __restrict__ sB, float2* __restrict__ slab,
                                            const Q* __restrict__ queue, const int qn, Q* __restrict__ rareq, int* rcount,
                                            Emit emit) {
    constexpr int NT = ST_THREADS;
    const int tid = threadIdx.x;
    for (int base = 0; base < qn; base += NT) {  // uniform trip count: the warp is converged at the ballot
        const int q = base + tid;
        const bool act = q < qn;
        const unsigned wm = __ballot_sync(0xffffffffu, act);
        if (act) {
            const unsigned e = queue[q];
            const int r = e >> SHIFT, c = e & ((1 << SHIFT) - 1);
            const float4* A = sA + r * REC_F4;
            const float4* B = sB + c * REC_F4;
            const float ov = overlap_area<FL>(A, B, slab + tid, NT, wm);
            if (ov < 0.f) rareq[atomicAdd(rcount, 1)] = (Q)e;  // at most qn <= capacity of the queues
            else emit(r, c, ov, A, B);
        }
    }
    __syncthreads();
    const int rn = *rcount;
    if (rn > 0) {
        if (tid < ST_SLOW_LANES) {
            // vertex slots 0..7 -> column tid, 8..15 -> column tid + 64; angles -> columns tid + 128 (as floats)
            auto slab16 = [&](int k) -> float2& { return slab[(k & 7) * NT + tid + ((k >> 3) << 6)]; };
            auto ang16 = [&](int k) -> float& { return reinterpret_cast<float*>(slab + (k >> 1) * NT + 128 + tid)[k & 1]; };
            for (int q = tid; q < rn; q += ST_SLOW_LANES) {
                const unsigned e = rareq[q];
                const int r = e >> SHIFT, c = e & ((1 << SHIFT) - 1);
                const float4* A = sA + r * REC_F4;
                const float4* B = sB + c * REC_F4;
                emit(r, c, overlap_area_slow<FL>(A, B, slab16, ang16), A, B);
            }
        }
        __syncthreads();
        if (tid == 0) *rcount = 0;
    }
}

}  // namespace lg
