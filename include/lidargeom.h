/*
 * lidargeom.h -- C ABI of liblidargeom.so: B200 (sm_100a) rotated 3D-box geometry ops.
 *
 * This is the drop-in boundary for the reference's two pybind11/torch extensions
 *   iou3d_nms_cuda        (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17, iou3d_nms.h:9-12)
 *   roiaware_pool3d_cuda  (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177; points_in_boxes_* only)
 * Each entry point below cites the reference function it replaces.  There are no torch / pybind
 * types in any signature: plain device pointers, sizes, and a CUDA stream passed as void*.
 *
 * Conventions
 *   - boxes are float32 rows of 7: (x, y, z, dx, dy, dz, heading), contiguous, 4-byte aligned
 *     (28-byte rows are NOT 16-byte aligned; the library never assumes they are).
 *   - points are float32 rows of 3: (x, y, z).
 *   - every pointer is a DEVICE pointer unless its name ends in _host.
 *   - the library never allocates, never synchronises and keeps no global state: outputs and the
 *     scratch `ws` are caller-owned (query the size with lg_*_workspace_bytes), work is enqueued on
 *     `stream` (a cudaStream_t; NULL = legacy default stream) and is stream-ordered.
 *   - return value: LG_OK (0), a positive cudaError_t from the launch, or a negative LG_ERR_* code.
 *     lg_last_error_string() describes the last failure on the calling thread.
 *     (The reference prints to stderr and calls exit(-1): iou3d_nms.cpp:14-38.)
 *   - empty inputs (n == 0, m == 0, ...) return LG_OK without launching anything.
 *
 * Arithmetic contract (DESIGN.md "arithmetic contract"): by default every IoU is computed with the
 * operation order AND the FMA contraction of the reference CUDA kernels as nvcc 12.9 builds them for
 * sm_100a, so results agree with the reference GPU path to the last bits (polygon-vertex ordering
 * ties aside).  LG_FLAG_STRICT_FP32 switches to the un-contracted arithmetic of the reference CPU
 * build (iou3d_cpu.cpp compiled by g++ -O2 on x86-64).
 */
#ifndef LIDARGEOM_H_
#define LIDARGEOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LG_API __attribute__((visibility("default")))
#else
#define LG_API
#endif

#define LG_VERSION 100 /* 0.1.0 */

#define LG_OK 0
#define LG_ERR_INVALID_ARG (-1)   /* null pointer, negative size, ld_out < m, ... */
#define LG_ERR_WORKSPACE (-2)     /* ws == NULL or ws_bytes too small */
#define LG_ERR_TOO_LARGE (-3)     /* size exceeds a documented limit */
#define LG_ERR_NO_DEVICE (-4)     /* no usable CUDA device / wrong architecture */

/* flags */
#define LG_FLAG_NONE 0u
#define LG_FLAG_STRICT_FP32 1u /* un-contracted FP32 (reference CPU build's rounding) instead of the reference CUDA build's */
#define LG_FLAG_NMS_NO_CLUSTER 4u /* lazy rotated NMS: one CTA per problem even when the batch is small enough to give every
                                    problem a thread-block cluster of 2-8 SMs (the default); same keep list either way */
#define LG_FLAG_IOU_SMALL_LIST 8u /* (testing) N x M IoU, two-phase sweep: shrink the survivor list to 1024 entries so that the
                                     overflow path (the complete one-kernel sweep on the flagged strips) is exercised */
#define LG_FLAG_IOU_ONE_KERNEL 16u /* N x M IoU: the complete one-kernel sweep whatever the size (the two-phase sweep is the default
                                      from 2^26 pairs on; results are identical) */
#define LG_FLAG_NMS_FULL_MASK 2u /* rotated NMS: materialise the reference's N x N/64 suppression mask (upper triangle) and sweep
                                    it, instead of the default lazy evaluation of kept rows only; same keep list either way */

/* limits */
#define LG_NMS_MAX_BOXES 262144 /* per NMS problem (the mask + sweep formulation then needs nmax^2 / 8 bytes of workspace per problem) */
#define LG_MAX_PEERS 8         /* ranks of one NVLink domain served by lg_nms_rotated_gather */
#define LG_PIB_MAX_BOXES 2048  /* boxes per frame for lg_points_in_boxes (records are shared-memory resident);
                                  frames of up to 254 boxes take the grid-culled path, larger ones test every box */
#define LG_ROIPOINT_MAX_SAMPLES 2048 /* sampled points per box for lg_roipoint_pool3d_forward (reference default: 512) */

LG_API int lg_version(void);
LG_API const char *lg_last_error_string(void);
/* 0 if a CUDA device of compute capability 10.x is current, else LG_ERR_NO_DEVICE. */
LG_API int lg_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * N x M rotated BEV overlap / IoU / 3D IoU.   out[i * ld_out + j], 64-bit offsets (the reference
 * indexes with 32-bit int and is wrong beyond 2^31 pairs: iou3d_nms_kernel.cu:248,264).
 *
 *   lg_boxes_overlap_bev  <- boxes_overlap_bev_gpu   (iou3d_nms.cpp:49-68,  kernel.cu:236-249)
 *   lg_boxes_iou_bev      <- boxes_iou_bev_gpu       (iou3d_nms.cpp:70-88,  kernel.cu:251-265)
 *   lg_boxes_iou3d        <- boxes_iou3d_gpu, the whole Python function fused into one pass
 *                            (iou3d_nms_utils.py:48-81: ~12 elementwise torch kernels + 6 temporaries)
 *
 * ws must hold lg_iou_workspace_bytes(n, m) bytes (per-box records).  Unlike the reference, `out`
 * does not have to be zero-filled beforehand: every element is written.
 */
LG_API size_t lg_iou_workspace_bytes(int64_t n, int64_t m);
LG_API int lg_boxes_overlap_bev(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, float *out, int64_t ld_out,
                         void *ws, size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_boxes_iou_bev(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, float *out, int64_t ld_out,
                     void *ws, size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_boxes_iou3d(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, float *out, int64_t ld_out,
                   void *ws, size_t ws_bytes, unsigned flags, void *stream);

/* Row / column maxima of the N x M matrix WITHOUT materialising it (the consumers' epilogue:
 * proposal_target_layer.py:107 `torch.max(iou3d, dim=1)`, axis_aligned_target_assigner.py:150-169 argmax over both axes,
 * detector3d_template.py:310-313; decisive at 200k x 200k, where the matrix is 160 GB).
 *   kind: 0 = overlap_bev, 1 = iou_bev, 2 = iou3d (same arithmetic as the matrix entry points)
 *   row_max[i] = max_j v(i, j), row_argmax[i] = the LOWEST j that attains it (torch.max's convention; 0 for an all-zero row);
 *   col_max / col_argmax likewise over i.  Any of the four output pointers may be NULL.
 * ws must hold lg_iou_reduce_workspace_bytes(n, m) bytes.  n, m < 2^32 - 1. */
LG_API size_t lg_iou_reduce_workspace_bytes(int64_t n, int64_t m);
LG_API int lg_boxes_iou_reduce(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, int kind, float *row_max,
                               int64_t *row_argmax, float *col_max, int64_t *col_argmax, void *ws, size_t ws_bytes, unsigned flags,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * NMS, rotated (nms_gpu: iou3d_nms.cpp:90-136, kernel.cu:267-311) and axis-aligned
 * (nms_normal_gpu: iou3d_nms.cpp:139-186, kernel.cu:314-372), batched over P independent problems.
 *
 * Problem p owns rows boxes[p * nmax .. p * nmax + counts[p]) (counts == NULL: all nmax rows).
 * The i-th box of a problem IN DESCENDING SCORE ORDER is
 *     order == NULL :  boxes[p*nmax + i]                       (caller already sorted and gathered)
 *     order != NULL :  boxes[p*nmax + order[p*nmax + i]]       (order = the wrapper's scores.sort()[1])
 * Box i is kept iff no kept k < i has iou(box_k, box_i) > thresh (strict, argument order (k, i)).
 * Output, entirely on the device (no host sweep, no D2H of the mask, no hidden sync):
 *     num_keep[p]            number of kept boxes
 *     keep[p*nmax + 0..num)  order == NULL: positions i; order != NULL: order[...] i.e. indices into
 *                            the caller's unsorted boxes -- exactly what the Python wrapper returns
 *                            (iou3d_nms_utils.py:99).  Entries beyond num_keep[p] are set to -1.
 * nmax <= LG_NMS_MAX_BOXES.
 *
 * Rotated NMS evaluates, by default, only the mask rows the sweep would read (rows of kept boxes): the
 * reference computes N x N IoUs per problem of which its sweep (iou3d_nms.cpp:121-132) consumes kept rows
 * only.  LG_FLAG_NMS_FULL_MASK selects the mask + sweep formulation; both give the identical keep list.
 */
LG_API size_t lg_nms_workspace_bytes(int num_problems, int nmax); /* enough for every variant and flag */
/* exact size for one variant: normal = 0 rotated / 1 axis-aligned; flags as passed to the NMS call */
LG_API size_t lg_nms_workspace_bytes_ex(int num_problems, int nmax, int normal, unsigned flags);
/* byte offset inside ws of three uint64 work counters the lazy rotated NMS leaves behind: [0] pairs put to the
 * exact-zero cull test, [1] pairs evaluated by the polygon path, [2] pairs with a non-zero overlap
 * (bench.py's roofline accounting reads them) */
LG_API size_t lg_nms_stats_offset(int num_problems, int nmax);
LG_API int lg_nms_rotated_batched(const float *boxes, const int64_t *order, const int32_t *counts, int num_problems, int nmax,
                           float thresh, void *ws, size_t ws_bytes, int64_t *keep, int32_t *num_keep, unsigned flags,
                           void *stream);
LG_API int lg_nms_normal_batched(const float *boxes, const int64_t *order, const int32_t *counts, int num_problems, int nmax,
                          float thresh, void *ws, size_t ws_bytes, int64_t *keep, int32_t *num_keep, unsigned flags,
                          void *stream);
/* The general form: `normal` 0 rotated / 1 axis-aligned; keep is (P, keep_ld) and only the first max_keep <= keep_ld entries
 * of a row are written (kept indices, then -1), num_keep[p] <= max_keep.  max_keep is the caller's NMS_POST_MAXSIZE
 * (model_nms_utils.py:20 `selected[:NMS_POST_MAXSIZE]`): the first max_keep kept boxes are exactly the reference's truncated
 * list, and the kernels stop choosing candidates once it is full.  The two entry points above are this one with
 * max_keep = keep_ld = nmax. */
LG_API int lg_nms_batched_ex(const float *boxes, const int64_t *order, const int32_t *counts, int num_problems, int nmax,
                             float thresh, int normal, int max_keep, int64_t keep_ld, void *ws, size_t ws_bytes, int64_t *keep,
                             int32_t *num_keep, unsigned flags, void *stream);
/* Rotated NMS fused with the gather of its results over NVLink peer memory (the reference merges per-rank results
 * through pickle files and two barriers, pcdet/utils/common_utils.py:206-227).  peer_bufs[r], r < num_peers <= LG_MAX_PEERS, is
 * the address IN THIS PROCESS of rank r's packed result buffer (rows, 1 + max_keep) int64 -- CUDA peer mappings, e.g. the
 * buffer_ptrs of a torch symmetric-memory allocation; the own rank's buffer is one of them.  Problem p's row row0 + p of EVERY
 * buffer receives (num_keep, kept indices ..., -1 ...), written by the NMS kernel's epilogue with plain stores; no keep tensor,
 * no collective.  The caller orders the ranks (a barrier after the call, before anyone reads; a buffer is not rewritten while a
 * peer may still read it: alternate two).  Lazy rotated NMS only (nmax up to ~12,000).  num_keep (P) may be NULL. */
LG_API int lg_nms_rotated_gather(const float *boxes, const int64_t *order, const int32_t *counts, int num_problems, int nmax,
                                 float thresh, int max_keep, void *ws, size_t ws_bytes, int64_t *const *peer_bufs, int num_peers,
                                 int64_t row0, int32_t *num_keep, unsigned flags, void *stream);
/* The same pipeline one phase at a time, for profiling (bench.py times the mask kernel alone for its
 * roofline line): phases is a bit-or of LG_NMS_PHASE_*; ws carries the records and the mask between calls.  The lazy rotated
 * NMS is ONE kernel (records included): it runs when LG_NMS_PHASE_SWEEP is set and ignores the other two bits. */
#define LG_NMS_PHASE_RECORDS 1u
#define LG_NMS_PHASE_MASK 2u
#define LG_NMS_PHASE_SWEEP 4u
LG_API int lg_nms_batched_phases(const float *boxes, const int64_t *order, const int32_t *counts, int num_problems, int nmax,
                                 float thresh, void *ws, size_t ws_bytes, int64_t *keep, int32_t *num_keep, unsigned flags,
                                 void *stream, int normal, unsigned phases);
/* single problem == batched with num_problems = 1, counts = NULL */
LG_API int lg_nms_rotated(const float *boxes, const int64_t *order, int n, float thresh, void *ws, size_t ws_bytes,
                   int64_t *keep, int32_t *num_keep, unsigned flags, void *stream);
LG_API int lg_nms_normal(const float *boxes, const int64_t *order, int n, float thresh, void *ws, size_t ws_bytes,
                  int64_t *keep, int32_t *num_keep, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * points in boxes.
 *   lg_points_in_boxes      <- points_in_boxes_gpu (roiaware_pool3d.cpp:98-118, kernel.cu:16-36,313-359)
 *       boxes (B, T, 7), pts (B, M, 3) -> out (B, M) int32: lowest box index containing the point,
 *       else -1 (written by the library; the reference needs the wrapper to pre-fill -1).
 *       MARGIN 1e-5 on x/y compared in double, z extent closed.  T <= LG_PIB_MAX_BOXES.
 *   lg_points_in_boxes_mask <- the all-pairs form of points_in_boxes_cpu (roiaware_pool3d.cpp:121-168)
 *       boxes (N, 7), pts (M, 3) -> out (N, M) int32 0/1 with the caller's margin (reference CPU: 1e-2).
 */
LG_API size_t lg_points_in_boxes_workspace_bytes(int batch, int num_boxes, int64_t num_points);
LG_API int lg_points_in_boxes(const float *boxes, const float *pts, int32_t *out, int batch, int num_boxes, int64_t num_points,
                       void *ws, size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_points_in_boxes_mask(const float *boxes, int64_t n, const float *pts, int64_t m, int32_t *out, float margin,
                            unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The other users of check_pt_in_box3d ("next" rows).
 *   lg_roiaware_pool3d_forward  <- roiaware_pool3d_gpu (roiaware_pool3d.cpp:25-62; launcher kernel.cu:186-226)
 *       rois (N, 7), pts (M, 3), pts_feature (M, C) -> pooled_features (N, ox, oy, oz, C) f32,
 *       argmax (N, ox, oy, oz, C) i32 (max pooling: point index or -1; may be NULL for avg pooling, else zero-filled),
 *       pts_idx_of_voxels (N, ox, oy, oz, max_pts) i32: [0] = count (<= max_pts - 1), then the voxel's first points in
 *       ascending index, zeros after.  pool_method 0 = max, 1 = avg.  Every output element is written by the library
 *       (the reference needs its wrapper to zero-fill all three, roiaware_pool3d_utils.py:84-86).  1 <= ox, oy, oz <= 255
 *       (the reference packs voxel coordinates into 8 bits each), ox*oy*oz <= 2^22.
 *   lg_roiaware_pool3d_backward <- roiaware_pool3d_gpu_backward (roiaware_pool3d.cpp:64-95; kernel.cu:229-310)
 *       ACCUMULATES into grad_in (num_pts, C) with float atomics like the reference: the caller zero-fills it
 *       (roiaware_pool3d_utils.py:104).  pts_idx_of_voxels / argmax are the forward's.
 *   lg_roipoint_pool3d_forward  <- roipool3d_gpu (roipoint_pool3d.cpp:24-58; kernel.cu:38-164)
 *       xyz (B, N, 3), boxes3d (B, M, 7) (already enlarged by the caller, roipoint_pool3d_utils.py:53), pts_feature (B, N, C)
 *       -> pooled_features (B, M, S, 3 + C): the first S inside points of every box in ascending index, repeated
 *       cyclically when fewer (zeros when none), pooled_empty_flag (B, M) i32 0/1.  Both fully written.
 */
LG_API int lg_roiaware_pool3d_forward(const float *rois, int num_rois, const float *pts, int num_pts, const float *pts_feature,
                                      int channels, int out_x, int out_y, int out_z, int max_pts_each_voxel, int pool_method,
                                      float *pooled_features, int32_t *argmax, int32_t *pts_idx_of_voxels, unsigned flags,
                                      void *stream);
LG_API int lg_roiaware_pool3d_backward(const int32_t *pts_idx_of_voxels, const int32_t *argmax, const float *grad_out, float *grad_in,
                                       int num_rois, int out_x, int out_y, int out_z, int channels, int max_pts_each_voxel,
                                       int pool_method, unsigned flags, void *stream);
LG_API int lg_roipoint_pool3d_forward(const float *xyz, const float *boxes3d, const float *pts_feature, int batch, int num_pts,
                                      int num_boxes, int channels, int num_sampled, float *pooled_features,
                                      int32_t *pooled_empty_flag, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * KITTI evaluation overlaps.  NOT the conventions of the ops above: 5-parameter boxes
 * (cx, cy, x_d, y_d, angle) with the angle clockwise-positive, no margin, closed containment, and
 * out[n * k_total + k] = f(query_boxes[k], boxes[n]); criterion -1: inter / (area_q + area_b - inter), 0: inter / area_q,
 * 1: inter / area_b, anything else: inter (rotate_iou.py:246-258).  Arithmetic: that of the reference's numba.cuda kernel
 * as numba / NVVM / ptxas 12.9 build it for sm_100a (float32 geometry, float64 area sum and ratio); LG_FLAG_STRICT_FP32
 * drops the FMA contraction.  The reference's 8-point polygon buffer (undefined behaviour beyond) is 24 points here.
 *
 *   lg_rotate_iou_eval      <- rotate_iou_gpu_eval + rotate_iou_kernel_eval (kitti_object_eval_python/rotate_iou.py:260-330);
 *                              bev_box_overlap (eval.py:111-113) is this call.  boxes (N, 5) f32, query_boxes (K, 5) f32.
 *   lg_d3_box_overlap       <- d3_box_overlap + d3_box_overlap_kernel (eval.py:116-155): float64 CAMERA boxes
 *                              (x, y, z, l, h, w, ry); BEV overlap of columns [0, 2, 3, 5, 6] cast to float32, height overlap
 *                              and ratio in float64, result float32 -- one pass instead of GPU kernel + D2H + CPU pass.
 *   lg_kitti_overlaps_parts <- the per-part loop of calculate_iou_partly (eval.py:340-395): num_parts independent
 *                              (gt part) x (dt part) problems in ONE launch.  gt_boxes (num_gt, 7) / dt_boxes (num_dt, 7)
 *                              float64 camera boxes of all parts end to end; gt_off / dt_off / out_off: (num_parts + 1)
 *                              int64 DEVICE arrays of row offsets and of output offsets (out_off[p + 1] - out_off[p] =
 *                              rows_gt(p) * rows_dt(p)); part p's matrix is out[out_off[p] ..) row-major (gt x dt).
 *                              metric 1 = bev (bev_box_overlap of columns [0,2,3,5,6]), 2 = 3d (d3_box_overlap).
 * ws must hold lg_kitti_workspace_bytes(rows of boxes, rows of query boxes, num_parts (0 for the single-problem calls)) bytes.
 */
LG_API size_t lg_kitti_workspace_bytes(int64_t num_boxes, int64_t num_query_boxes, int num_parts);
LG_API int lg_rotate_iou_eval(const float *boxes, int64_t n, const float *query_boxes, int64_t k, float *out, int criterion, void *ws,
                              size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_d3_box_overlap(const double *boxes, int64_t n, const double *qboxes, int64_t k, float *out, int criterion, void *ws,
                             size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_kitti_overlaps_parts(const double *gt_boxes, int64_t num_gt, const double *dt_boxes, int64_t num_dt,
                                   const int64_t *gt_off, const int64_t *dt_off, const int64_t *out_off, int num_parts,
                                   int64_t num_out, int metric, int criterion, float *out, void *ws, size_t ws_bytes, unsigned flags,
                                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * Selection steps of the post-processing front end (pcdet/models/model_utils/model_nms_utils.py:6-25,
 * called per frame / per class from detector3d_template.py:190-260), batched over P problems.
 *   lg_select_topk   <- `scores >= SCORE_THRESH` mask, torch.topk(k = NMS_PRE_MAXSIZE), the box gather:
 *       scores (P, n) f32; problem p takes its boxes from frame p / problems_per_frame:
 *       row i of that frame = boxes + frame * box_frame_stride + i * box_row_stride (strides in floats; >= 7 floats per row)
 *       -> top_idx (P, k) int64: candidate indices in descending score, equal scores by ascending index (0 beyond counts[p]);
 *          counts (P) int32 = min(k, candidates passing the threshold); top_boxes (P, k, 7) f32 (0 beyond counts[p]; may be NULL).
 *       use_thresh = 0: every candidate passes.  k <= LG_SELECT_MAX_K, n < 2^31.
 *       ws: lg_select_workspace_bytes(P, n) bytes.
 *   lg_select_finish <- keep[:NMS_POST_MAXSIZE], indices[keep], scores[selected] (model_nms_utils.py:21-25):
 *       keep (P, k) / num_keep (P) as lg_nms_*_batched return them for the top_boxes above
 *       -> selected (P, post) int64 candidate indices (-1 padded), num_out (P) = min(num_keep, post), sel_scores (P, post).
 */
#define LG_SELECT_MAX_K 4096
LG_API size_t lg_select_workspace_bytes(int num_problems, int64_t n);
LG_API int lg_select_topk(const float *scores, int num_problems, int64_t n, int k, float score_thresh, int use_thresh, const float *boxes,
                          int64_t box_frame_stride, int64_t box_row_stride, int problems_per_frame, int64_t *top_idx, int32_t *counts,
                          float *top_boxes, void *ws, size_t ws_bytes, unsigned flags, void *stream);
LG_API int lg_select_finish(const int64_t *keep, const int32_t *num_keep, const int64_t *top_idx, const float *scores, int num_problems,
                            int64_t n, int k, int post, int64_t *selected, int32_t *num_out, float *sel_scores, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDARGEOM_H_ */
