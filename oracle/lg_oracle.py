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
