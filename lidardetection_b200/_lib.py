"""ctypes binding of liblidargeom.so (include/lidargeom.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LG_LIB_PATH") or os.path.join(_HERE, "liblidargeom.so")  # LG_LIB_PATH: developer builds (tools/lz_timing.py)

LG_FLAG_NONE = 0
LG_FLAG_STRICT_FP32 = 1
LG_FLAG_NMS_FULL_MASK = 2
LG_FLAG_NMS_NO_CLUSTER = 4
LG_FLAG_IOU_SMALL_LIST = 8
LG_FLAG_IOU_ONE_KERNEL = 16
LG_NMS_MAX_BOXES = 262144
LG_PIB_MAX_BOXES = 2048

_lib = None


class LidarGeomError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python lidardetection_b200/csrc/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for these ops."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, sz, u32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_uint, C.c_float
    L.lg_version.restype = C.c_int
    L.lg_last_error_string.restype = C.c_char_p
    L.lg_check_device.restype = C.c_int
    L.lg_iou_workspace_bytes.restype = sz
    L.lg_iou_workspace_bytes.argtypes = [i64, i64]
    for name in ("lg_boxes_overlap_bev", "lg_boxes_iou_bev", "lg_boxes_iou3d"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, i64, vp, i64, vp, i64, vp, sz, u32, vp]
    L.lg_iou_reduce_workspace_bytes.restype = sz
    L.lg_iou_reduce_workspace_bytes.argtypes = [i64, i64]
    L.lg_boxes_iou_reduce.restype = C.c_int
    L.lg_boxes_iou_reduce.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, vp, vp, sz, u32, vp]
    L.lg_nms_workspace_bytes.restype = sz
    L.lg_nms_workspace_bytes.argtypes = [i32, i32]
    L.lg_nms_workspace_bytes_ex.restype = sz
    L.lg_nms_workspace_bytes_ex.argtypes = [i32, i32, i32, u32]
    L.lg_nms_stats_offset.restype = sz
    L.lg_nms_stats_offset.argtypes = [i32, i32]
    for name in ("lg_nms_rotated_batched", "lg_nms_normal_batched"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, vp, i32, i32, f32, vp, sz, vp, vp, u32, vp]
    L.lg_nms_batched_ex.restype = C.c_int
    L.lg_nms_batched_ex.argtypes = [vp, vp, vp, i32, i32, f32, i32, i32, i64, vp, sz, vp, vp, u32, vp]
    L.lg_nms_rotated_gather.restype = C.c_int
    L.lg_nms_rotated_gather.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp, sz, C.POINTER(C.c_void_p), i32, i64, vp, u32, vp]
    L.lg_nms_batched_phases.restype = C.c_int
    L.lg_nms_batched_phases.argtypes = [vp, vp, vp, i32, i32, f32, vp, sz, vp, vp, u32, vp, i32, u32]
    for name in ("lg_nms_rotated", "lg_nms_normal"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, i32, f32, vp, sz, vp, vp, u32, vp]
    L.lg_points_in_boxes_workspace_bytes.restype = sz
    L.lg_points_in_boxes_workspace_bytes.argtypes = [i32, i32, i64]
    L.lg_points_in_boxes.restype = C.c_int
    L.lg_points_in_boxes.argtypes = [vp, vp, vp, i32, i32, i64, vp, sz, u32, vp]
    L.lg_points_in_boxes_mask.restype = C.c_int
    L.lg_points_in_boxes_mask.argtypes = [vp, i64, vp, i64, vp, f32, u32, vp]
    L.lg_roiaware_pool3d_forward.restype = C.c_int
    L.lg_roiaware_pool3d_forward.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, u32, vp]
    L.lg_roiaware_pool3d_backward.restype = C.c_int
    L.lg_roiaware_pool3d_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, u32, vp]
    L.lg_roipoint_pool3d_forward.restype = C.c_int
    L.lg_roipoint_pool3d_forward.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, u32, vp]
    L.lg_kitti_workspace_bytes.restype = sz
    L.lg_kitti_workspace_bytes.argtypes = [i64, i64, i32]
    L.lg_rotate_iou_eval.restype = C.c_int
    L.lg_rotate_iou_eval.argtypes = [vp, i64, vp, i64, vp, i32, vp, sz, u32, vp]
    L.lg_d3_box_overlap.restype = C.c_int
    L.lg_d3_box_overlap.argtypes = [vp, i64, vp, i64, vp, i32, vp, sz, u32, vp]
    L.lg_kitti_overlaps_parts.restype = C.c_int
    L.lg_kitti_overlaps_parts.argtypes = [vp, i64, vp, i64, vp, vp, vp, i32, i64, i32, i32, vp, vp, sz, u32, vp]
    L.lg_select_workspace_bytes.restype = sz
    L.lg_select_workspace_bytes.argtypes = [i32, i64]
    L.lg_select_topk.restype = C.c_int
    L.lg_select_topk.argtypes = [vp, i32, i64, i32, f32, i32, vp, i64, i64, i32, vp, vp, vp, vp, sz, u32, vp]
    L.lg_select_finish.restype = C.c_int
    L.lg_select_finish.argtypes = [vp, vp, vp, vp, i32, i64, i32, i32, vp, vp, vp, vp]
    _lib = L
    return L


EXPORTS = [
    "lg_version", "lg_last_error_string", "lg_check_device",
    "lg_iou_workspace_bytes", "lg_boxes_overlap_bev", "lg_boxes_iou_bev", "lg_boxes_iou3d", "lg_iou_reduce_workspace_bytes", "lg_boxes_iou_reduce",
    "lg_nms_workspace_bytes", "lg_nms_workspace_bytes_ex", "lg_nms_stats_offset", "lg_nms_rotated_batched", "lg_nms_normal_batched", "lg_nms_batched_ex", "lg_nms_rotated_gather", "lg_nms_batched_phases", "lg_nms_rotated", "lg_nms_normal",
    "lg_points_in_boxes_workspace_bytes", "lg_points_in_boxes", "lg_points_in_boxes_mask",
    "lg_roiaware_pool3d_forward", "lg_roiaware_pool3d_backward", "lg_roipoint_pool3d_forward",
    "lg_kitti_workspace_bytes", "lg_rotate_iou_eval", "lg_d3_box_overlap", "lg_kitti_overlaps_parts",
    "lg_select_workspace_bytes", "lg_select_topk", "lg_select_finish",
]


def check(rc, what):
    if rc != 0:
        msg = lib().lg_last_error_string().decode("utf-8", "replace")
        raise LidarGeomError(f"{what} failed with status {rc}: {msg}")


def require_usable_cuda(what):
    """The *_cpu entry points of the reference API are served by the GPU (no CPU fallback by contract).  The reference calls
    them from dataset code that may run in forked DataLoader workers (database_sampler.py:212-216, box_utils.py:85); CUDA
    cannot be initialised in a process forked from one that already holds a context, so say so instead of crashing inside
    the driver ("DataLoader workers")."""
    import torch

    if torch.cuda._is_in_bad_fork():
        raise LidarGeomError(
            f"{what}: this process was forked from a parent that had already initialised CUDA, so it cannot use the GPU.  {what} runs on "
            "the GPU here (there is no CPU fallback).  Start the workers with the 'spawn' or 'forkserver' method -- "
            "DataLoader(..., multiprocessing_context='spawn') -- or call it with num_workers=0; see INTEGRATION.md.")
    if not torch.cuda.is_available():
        raise LidarGeomError(f"{what}: no CUDA device is visible; {what} runs on the GPU here (there is no CPU fallback)")


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
