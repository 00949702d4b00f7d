"""ctypes binding of oracle/liblg_oracle.so (numpy in / numpy out).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import build as _build

FLAVOR_CPU = 0   # reference CPU build: g++ -O2, no FMA contraction, glibc trig (iou3d_cpu.cpp)
FLAVOR_CUDA = 1  # reference CUDA build: nvcc 12.9 sm_100a contraction + libdevice sinf/cosf

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblg_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            _build.build()
        L = C.CDLL(path)
        fp, ip, i64p, u64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64))
        L.lgo_sinf.restype = C.c_float
        L.lgo_sinf.argtypes = [C.c_float, C.c_int]
        L.lgo_cosf.restype = C.c_float
        L.lgo_cosf.argtypes = [C.c_float, C.c_int]
        for name in ("lgo_box_overlap", "lgo_iou_bev", "lgo_iou3d_pair"):
            f = getattr(L, name)
            f.restype = C.c_float
            f.argtypes = [fp, fp, C.c_int]
        L.lgo_box_overlap_cnt.restype = C.c_int
        L.lgo_box_overlap_cnt.argtypes = [fp, fp, C.c_int]
        L.lgo_iou_normal.restype = C.c_float
        L.lgo_iou_normal.argtypes = [fp, fp, C.c_int]
        for name in ("lgo_boxes_overlap_bev", "lgo_boxes_iou_bev", "lgo_boxes_iou3d"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [fp, C.c_int64, fp, C.c_int64, fp, C.c_int64, C.c_int]
        L.lgo_nms_mask.restype = None
        L.lgo_nms_mask.argtypes = [fp, C.c_int, C.c_float, C.c_int, C.c_int, u64p]
        L.lgo_nms_sweep.restype = C.c_int
        L.lgo_nms_sweep.argtypes = [u64p, C.c_int, i64p]
        L.lgo_nms.restype = C.c_int
        L.lgo_nms.argtypes = [fp, C.c_int, C.c_float, C.c_int, C.c_int, i64p]
        L.lgo_points_in_boxes_idx.restype = None
        L.lgo_points_in_boxes_idx.argtypes = [fp, fp, ip, C.c_int, C.c_int, C.c_int64, C.c_int]
        L.lgo_points_in_boxes_mask.restype = None
        L.lgo_points_in_boxes_mask.argtypes = [fp, C.c_int64, fp, C.c_int64, ip, C.c_float, C.c_int]
        L.lgo_point_in_box.restype = C.c_int
        L.lgo_point_in_box.argtypes = [fp, fp, C.c_float, C.c_int]
        L.lgo_roiaware_pool3d_forward.restype = None
        L.lgo_roiaware_pool3d_forward.argtypes = [fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  fp, ip, ip, C.c_int]
        L.lgo_roiaware_pool3d_backward.restype = None
        L.lgo_roiaware_pool3d_backward.argtypes = [ip, ip, fp, fp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.lgo_roipoint_pool3d_forward.restype = None
        L.lgo_roipoint_pool3d_forward.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, ip, C.c_int]
        dp = C.POINTER(C.c_double)
        L.lgo_rotate_iou_eval_pair.restype = C.c_float
        L.lgo_rotate_iou_eval_pair.argtypes = [fp, fp, C.c_int, C.c_int]
        L.lgo_rotate_iou_eval_cnt.restype = C.c_int
        L.lgo_rotate_iou_eval_cnt.argtypes = [fp, fp, C.c_int]
        L.lgo_rotate_iou_eval.restype = None
        L.lgo_rotate_iou_eval.argtypes = [fp, C.c_int64, fp, C.c_int64, fp, C.c_int, C.c_int]
        L.lgo_d3_box_overlap.restype = None
        L.lgo_d3_box_overlap.argtypes = [dp, C.c_int64, dp, C.c_int64, fp, C.c_int, C.c_int]
        _lib = L
    return _lib


def _f32(a, cols):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    assert a.ndim >= 2 and a.shape[-1] == cols, a.shape
    return a


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def sinf(x, flavor):
    return lib().lgo_sinf(float(np.float32(x)), flavor)


def cosf(x, flavor):
    return lib().lgo_cosf(float(np.float32(x)), flavor)


def _nm(fn, a, b, flavor):
    a, b = _f32(a, 7), _f32(b, 7)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    if out.size:
        getattr(lib(), fn)(_p(a, C.c_float), a.shape[0], _p(b, C.c_float), b.shape[0], _p(out, C.c_float), b.shape[0], flavor)
    return out


def boxes_overlap_bev(a, b, flavor=FLAVOR_CUDA):
    return _nm("lgo_boxes_overlap_bev", a, b, flavor)


def boxes_iou_bev(a, b, flavor=FLAVOR_CUDA):
    return _nm("lgo_boxes_iou_bev", a, b, flavor)


def boxes_iou3d(a, b, flavor=FLAVOR_CUDA):
    return _nm("lgo_boxes_iou3d", a, b, flavor)


def iou_bev_pair(a, b, flavor=FLAVOR_CUDA):
    a, b = _f32(np.reshape(a, (1, 7)), 7), _f32(np.reshape(b, (1, 7)), 7)
    return lib().lgo_iou_bev(_p(a, C.c_float), _p(b, C.c_float), flavor)


def iou_normal_pair(a, b, flavor=FLAVOR_CUDA):
    a, b = _f32(np.reshape(a, (1, 7)), 7), _f32(np.reshape(b, (1, 7)), 7)
    return lib().lgo_iou_normal(_p(a, C.c_float), _p(b, C.c_float), flavor)


def overlap_cnt_pair(a, b, flavor=FLAVOR_CUDA):
    a, b = _f32(np.reshape(a, (1, 7)), 7), _f32(np.reshape(b, (1, 7)), 7)
    return lib().lgo_box_overlap_cnt(_p(a, C.c_float), _p(b, C.c_float), flavor)


def nms_mask(boxes_sorted, thresh, normal=False, flavor=FLAVOR_CUDA):
    b = _f32(boxes_sorted, 7)
    n = b.shape[0]
    mask = np.zeros((n, (n + 63) // 64), dtype=np.uint64)
    if n:
        lib().lgo_nms_mask(_p(b, C.c_float), n, thresh, int(normal), flavor, _p(mask, C.c_uint64))
    return mask


def nms_sweep(mask):
    mask = np.ascontiguousarray(mask, dtype=np.uint64)
    n = mask.shape[0]
    keep = np.zeros(max(n, 1), dtype=np.int64)
    num = lib().lgo_nms_sweep(_p(mask, C.c_uint64), n, _p(keep, C.c_int64)) if n else 0
    return keep[:num].copy()


def nms_sorted(boxes_sorted, thresh, normal=False, flavor=FLAVOR_CUDA, lazy=True):
    """keep positions (into the score-sorted boxes).  lazy=False goes through the full mask + sweep."""
    b = _f32(boxes_sorted, 7)
    n = b.shape[0]
    if not lazy:
        return nms_sweep(nms_mask(b, thresh, normal, flavor))
    keep = np.zeros(max(n, 1), dtype=np.int64)
    num = lib().lgo_nms(_p(b, C.c_float), n, thresh, int(normal), flavor, _p(keep, C.c_int64)) if n else 0
    return keep[:num].copy()


def nms(boxes, scores, thresh, pre_maxsize=None, normal=False, flavor=FLAVOR_CUDA, order=None):
    """Restates the wrapper (iou3d_nms_utils.py:84-116): indices into the ORIGINAL boxes, score order.
    `order` may be supplied to pin the (unstable) torch sort permutation."""
    boxes = _f32(boxes, 7)
    if order is None:
        order = np.argsort(-np.asarray(scores, dtype=np.float32), kind="stable")
    order = np.asarray(order, dtype=np.int64)
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = nms_sorted(boxes[order], thresh, normal, flavor)
    return order[keep]


def points_in_boxes_idx(points, boxes, flavor=FLAVOR_CUDA):
    """GPU-form semantics: points (B,M,3), boxes (B,T,7) -> (B,M) int32, first hit or -1, margin 1e-5."""
    pts, bx = _f32(points, 3), _f32(boxes, 7)
    assert pts.ndim == 3 and bx.ndim == 3 and pts.shape[0] == bx.shape[0]
    B, M, _ = pts.shape
    T = bx.shape[1]
    out = np.full((B, M), -1, dtype=np.int32)
    if out.size:
        lib().lgo_points_in_boxes_idx(_p(bx, C.c_float), _p(pts, C.c_float), _p(out, C.c_int32), B, T, M, flavor)
    return out


def points_in_boxes_mask(points, boxes, margin=1e-2, flavor=FLAVOR_CPU):
    """CPU-form semantics: points (M,3), boxes (N,7) -> (N,M) int32 0/1, margin 1e-2."""
    pts, bx = _f32(points, 3), _f32(boxes, 7)
    out = np.zeros((bx.shape[0], pts.shape[0]), dtype=np.int32)
    if out.size:
        lib().lgo_points_in_boxes_mask(_p(bx, C.c_float), bx.shape[0], _p(pts, C.c_float), pts.shape[0], _p(out, C.c_int32), margin, flavor)
    return out


def roiaware_pool3d_forward(rois, pts, pts_feature, out_size, max_pts_each_voxel=128, pool_method="max", flavor=FLAVOR_CUDA):
    """RoIAwarePool3dFunction.forward (roiaware_pool3d_utils.py:57-93): rois (N,7), pts (M,3), pts_feature (M,C) ->
    pooled (N,ox,oy,oz,C) f32, argmax (N,ox,oy,oz,C) i32, pts_idx_of_voxels (N,ox,oy,oz,max_pts) i32."""
    rois, pts = _f32(rois, 7), _f32(pts, 3)
    feat = np.ascontiguousarray(np.asarray(pts_feature, dtype=np.float32))
    ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else out_size
    n, m, c = rois.shape[0], pts.shape[0], feat.shape[1]
    pooled = np.zeros((n, ox, oy, oz, c), dtype=np.float32)
    argmax = np.zeros((n, ox, oy, oz, c), dtype=np.int32)
    pts_idx = np.zeros((n, ox, oy, oz, max_pts_each_voxel), dtype=np.int32)
    if n:
        lib().lgo_roiaware_pool3d_forward(_p(rois, C.c_float), n, _p(pts, C.c_float), m, _p(feat, C.c_float), c, ox, oy, oz,
                                          max_pts_each_voxel, {"max": 0, "avg": 1}[pool_method], _p(pooled, C.c_float),
                                          _p(argmax, C.c_int32), _p(pts_idx, C.c_int32), flavor)
    return pooled, argmax, pts_idx


def roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, num_pts, pool_method="max"):
    """RoIAwarePool3dFunction.backward (roiaware_pool3d_utils.py:95-107) -> grad_in (num_pts, C)."""
    pts_idx = np.ascontiguousarray(pts_idx_of_voxels, dtype=np.int32)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    g = np.ascontiguousarray(grad_out, dtype=np.float32)
    n, c, max_pts = pts_idx.shape[0], g.shape[-1], pts_idx.shape[-1]
    V = int(np.prod(pts_idx.shape[1:4]))
    grad_in = np.zeros((num_pts, c), dtype=np.float32)
    if n:
        lib().lgo_roiaware_pool3d_backward(_p(pts_idx, C.c_int32), _p(argmax, C.c_int32), _p(g, C.c_float), _p(grad_in, C.c_float),
                                           n, V, c, max_pts, {"max": 0, "avg": 1}[pool_method])
    return grad_in


def roipoint_pool3d_forward(points, point_features, boxes3d, num_sampled_points=512, flavor=FLAVOR_CUDA):
    """roipool3d_gpu (roipoint_pool3d.cpp:24-58) on already enlarged boxes: points (B,N,3), point_features (B,N,C),
    boxes3d (B,M,7) -> pooled (B,M,S,3+C) f32, empty flag (B,M) i32."""
    xyz, bx = _f32(points, 3), _f32(boxes3d, 7)
    feat = np.ascontiguousarray(np.asarray(point_features, dtype=np.float32))
    B, n, _ = xyz.shape
    m, c = bx.shape[1], feat.shape[2]
    pooled = np.zeros((B, m, num_sampled_points, 3 + c), dtype=np.float32)
    flag = np.zeros((B, m), dtype=np.int32)
    if B and m:
        lib().lgo_roipoint_pool3d_forward(_p(xyz, C.c_float), _p(bx, C.c_float), _p(feat, C.c_float), B, n, m, c,
                                          num_sampled_points, _p(pooled, C.c_float), _p(flag, C.c_int32), flavor)
    return pooled, flag


# ---- "next" row 8f-2: KITTI-eval rotated IoU (kitti_object_eval_python/rotate_iou.py, eval.py:111-155) ----
def rotate_iou_eval(boxes, query_boxes, criterion=-1, flavor=FLAVOR_CUDA):
    """(N,5), (K,5) -> (N,K) float32; iou[n,k] = devRotateIoUEval(query_boxes[k], boxes[n], criterion)"""
    b, q = _f32(np.asarray(boxes).reshape(-1, 5), 5), _f32(np.asarray(query_boxes).reshape(-1, 5), 5)
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if out.size:
        lib().lgo_rotate_iou_eval(_p(b, C.c_float), b.shape[0], _p(q, C.c_float), q.shape[0], _p(out, C.c_float), int(criterion), flavor)
    return out


def rotate_iou_eval_cnt(boxes, query_boxes, flavor=FLAVOR_CUDA):
    """number of polygon points the reference collects per pair ((N,K) int32); > 8 overflows the reference's buffer"""
    b, q = _f32(np.asarray(boxes).reshape(-1, 5), 5), _f32(np.asarray(query_boxes).reshape(-1, 5), 5)
    out = np.zeros((b.shape[0], q.shape[0]), np.int32)
    for i in range(b.shape[0]):
        for j in range(q.shape[0]):
            out[i, j] = lib().lgo_rotate_iou_eval_cnt(_p(q[j], C.c_float), _p(b[i], C.c_float), flavor)
    return out


def d3_box_overlap(boxes, qboxes, criterion=-1, flavor=FLAVOR_CUDA):
    """float64 camera boxes (N,7), (K,7) -> (N,K) float32 (eval.py:150-155)"""
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 7))
    q = np.ascontiguousarray(np.asarray(qboxes, dtype=np.float64).reshape(-1, 7))
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if out.size:
        lib().lgo_d3_box_overlap(_p(b, C.c_double), b.shape[0], _p(q, C.c_double), q.shape[0], _p(out, C.c_float), int(criterion), flavor)
    return out


def bev_box_overlap(boxes, qboxes, criterion=-1, flavor=FLAVOR_CUDA):
    return rotate_iou_eval(boxes, qboxes, criterion, flavor)
