"""Drop-in for the points_in_boxes_* part of pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41.

RoIAwarePool3d (the voxel pooling layer of the same reference module) is outside the hot path named by
BASELINE.json and is not provided here ("next").
"""
import torch

from ... import _lib
from ...utils import common_utils


def points_in_boxes_cpu(points, boxes):
    """roiaware_pool3d_utils.py:9-25.  numpy / CPU tensors in, same kind out, MARGIN = 1e-2.

    The reference loops boxes x points on one host thread (roiaware_pool3d.cpp:143-168); here the
    (N, num_points) mask is produced on the B200 (un-contracted arithmetic of the CPU build) and
    copied back.  There is no host fallback.
    Args:
        points: (num_points, 3)
        boxes: [x, y, z, dx, dy, dz, heading], (x, y, z) is the box center, each box DO NOT overlaps
    Returns:
        point_indices: (N, num_points)
    """
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = common_utils.check_numpy_to_torch(points)
    boxes, is_numpy = common_utils.check_numpy_to_torch(boxes)
    dev = torch.device('cuda', torch.cuda.current_device())
    out = points_in_boxes_mask_gpu(points.float().to(dev), boxes.float().to(dev), margin=1e-2,
                                   flags=_lib.LG_FLAG_STRICT_FP32)
    point_indices = out.cpu()
    return point_indices.numpy() if is_numpy else point_indices


def points_in_boxes_mask_gpu(points, boxes, margin=1e-2, flags=_lib.LG_FLAG_NONE):
    """All-pairs form on the device: points (M, 3), boxes (N, 7) cuda -> (N, M) int32 0/1."""
    assert points.is_cuda and boxes.is_cuda
    p, b = points.contiguous().float(), boxes.contiguous().float()
    n, m = b.shape[0], p.shape[0]
    out = torch.empty((n, m), dtype=torch.int32, device=p.device)
    if n == 0 or m == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(p.device):
        rc = L.lg_points_in_boxes_mask(_lib.ptr(b), n, _lib.ptr(p), m, _lib.ptr(out), float(margin), flags,
                                       _lib.stream_ptr(p.device))
    _lib.check(rc, 'lg_points_in_boxes_mask')
    return out


def points_in_boxes_gpu(points, boxes):
    """roiaware_pool3d_utils.py:28-41.
    :param points: (B, M, 3)
    :param boxes: (B, T, 7), num_valid_boxes <= T
    :return box_idxs_of_pts: (B, M), default background = -1
    """
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    batch_size, num_points, _ = points.shape
    assert points.is_cuda and boxes.is_cuda
    p, b = points.contiguous().float(), boxes.contiguous().float()
    box_idxs_of_pts = torch.empty((batch_size, num_points), dtype=torch.int, device=p.device)
    if batch_size == 0 or num_points == 0:
        return box_idxs_of_pts
    L = _lib.lib()
    with torch.cuda.device(p.device):
        rc = L.lg_points_in_boxes(_lib.ptr(b), _lib.ptr(p), _lib.ptr(box_idxs_of_pts), batch_size, b.shape[1], num_points,
                                  None, 0, _lib.LG_FLAG_NONE, _lib.stream_ptr(p.device))
    _lib.check(rc, 'lg_points_in_boxes')
    return box_idxs_of_pts
