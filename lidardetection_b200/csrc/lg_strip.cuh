// lg_strip.cuh -- the cull -> compact -> drain machinery shared by the IoU and NMS-mask kernels.
//
// A CTA (256 threads) owns a strip of the pair space: ST_ROWS rows x up to ST_COLS columns whose box
// records sit in shared memory.  It sweeps the strip in 64 x 64 tiles:
//   cull    lanes along columns, row records broadcast from smem; (ca-cb)^2 > (ra+rb)^2 proves the
//           reference result is exactly +0.0.  The 16 row iterations of a warp are fully unrolled, their
//           ballots kept in registers, so a warp does ONE shared atomicAdd per tile to reserve queue space;
//   queue   survivors of ALL tiles of the strip accumulate in one smem queue (16-bit (row, col) codes);
//   drain   every thread takes queue entries round-robin, so the expensive polygon path runs with full
//           warps whatever the survivor density (2-6 % for NMS, 0.3 % for anchors x GT, 100 % dense).
// The queue is drained when it could overflow on the next tile, and at the end of the strip.
#pragma once
#include "lg_geom.cuh"

namespace lg {

constexpr int ST_THREADS = 256;
constexpr int ST_TILE = 64;
constexpr int ST_ROWS = 64;
constexpr int ST_MAXT = 8;                  // column tiles per strip
constexpr int ST_COLS = ST_TILE * ST_MAXT;  // 512
constexpr int ST_QCAP = 8192;               // queue entries (u16 codes: row << 9 | col)

struct StripSmem {
    static constexpr size_t a_bytes = (size_t)ST_ROWS * REC_F4 * sizeof(float4);
    static constexpr size_t b_bytes = (size_t)ST_COLS * REC_F4 * sizeof(float4);
    static constexpr size_t slab_bytes = (size_t)16 * ST_THREADS * sizeof(float2);
    static constexpr size_t queue_bytes = (size_t)ST_QCAP * sizeof(uint16_t);
    static constexpr size_t extra_bytes = (size_t)ST_ROWS * ST_MAXT * 2 * sizeof(unsigned int);  // NMS mask words
    static constexpr size_t total = a_bytes + b_bytes + slab_bytes + queue_bytes + extra_bytes;
};

// Reserve queue space for this warp's survivors of one tile and write their codes.
// m[k] = ballot of the k-th row iteration (row = rbase + 4k), code = row << SHIFT | col.
template <int SHIFT>
__device__ __forceinline__ void push_survivors(const unsigned (&m)[16], const int lane, const int rbase, const int rstep,
                                               const int col, int* qcount, uint16_t* __restrict__ queue) {
    int total = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) total += __popc(m[k]);
    if (total == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(qcount, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (m[k]) {
            if ((m[k] >> lane) & 1u) queue[base + __popc(m[k] & lt)] = (uint16_t)(((rbase + rstep * k) << SHIFT) | col);
            base += __popc(m[k]);
        }
    }
}

}  // namespace lg
