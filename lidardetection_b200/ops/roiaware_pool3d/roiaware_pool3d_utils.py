"""Drop-in for pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py: points_in_boxes_cpu / points_in_boxes_gpu (:9-41, the hot
path named by BASELINE.json) and RoIAwarePool3d / RoIAwarePool3dFunction (:44-107, SURVEY.md section 8f-3, a "next" row).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

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
    _lib.require_usable_cuda('points_in_boxes_cpu')
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
    T = b.shape[1]

    def call(boxes_chunk, out):
        with torch.cuda.device(p.device):
            rc = L.lg_points_in_boxes(_lib.ptr(boxes_chunk), _lib.ptr(p), _lib.ptr(out), batch_size, boxes_chunk.shape[1], num_points,
                                      None, 0, _lib.LG_FLAG_NONE, _lib.stream_ptr(p.device))
        _lib.check(rc, 'lg_points_in_boxes')

    if T <= _lib.LG_PIB_MAX_BOXES:
        call(b, box_idxs_of_pts)
        return box_idxs_of_pts
    # more boxes per frame than the kernel keeps in shared memory (the reference has no limit): chunks of LG_PIB_MAX_BOXES, the
    # first (lowest-index) hit wins, as in the reference's `break` (roiaware_pool3d_kernel.cu:328-334)
    box_idxs_of_pts.fill_(-1)
    part = torch.empty_like(box_idxs_of_pts)
    for t0 in range(0, T, _lib.LG_PIB_MAX_BOXES):
        call(b[:, t0:t0 + _lib.LG_PIB_MAX_BOXES].contiguous(), part)
        take = (box_idxs_of_pts < 0) & (part >= 0)
        box_idxs_of_pts = torch.where(take, part + t0, box_idxs_of_pts)
    return box_idxs_of_pts


class RoIAwarePool3d(nn.Module):
    """roiaware_pool3d_utils.py:44-53."""

    def __init__(self, out_size, max_pts_each_voxel=128):
        super().__init__()
        self.out_size = out_size
        self.max_pts_each_voxel = max_pts_each_voxel

    def forward(self, rois, pts, pts_feature, pool_method='max'):
        assert pool_method in ['max', 'avg']
        return RoIAwarePool3dFunction.apply(rois, pts, pts_feature, self.out_size, self.max_pts_each_voxel, pool_method)


class RoIAwarePool3dFunction(Function):
    """roiaware_pool3d_utils.py:56-107 over lg_roiaware_pool3d_forward / _backward."""

    @staticmethod
    def forward(ctx, rois, pts, pts_feature, out_size, max_pts_each_voxel, pool_method):
        """
        Args:
            rois: (N, 7) [x, y, z, dx, dy, dz, heading] (x, y, z) is the box center
            pts: (npoints, 3)
            pts_feature: (npoints, C)
            out_size: int or tuple, like 7 or (7, 7, 7)
            max_pts_each_voxel:
            pool_method: 'max' or 'avg'
        Returns:
            pooled_features: (N, out_x, out_y, out_z, C)
        """
        assert rois.shape[1] == 7 and pts.shape[1] == 3
        if isinstance(out_size, int):
            out_x = out_y = out_z = out_size
        else:
            assert len(out_size) == 3
            for k in range(3):
                assert isinstance(out_size[k], int)
            out_x, out_y, out_z = out_size
        pooled_features, argmax, pts_idx_of_voxels = roiaware_pool3d_forward(
            rois, pts, pts_feature, (out_x, out_y, out_z), max_pts_each_voxel, pool_method)
        pool_method = {'max': 0, 'avg': 1}[pool_method]
        ctx.roiaware_pool3d_for_backward = (pts_idx_of_voxels, argmax, pool_method, pts.shape[0], pts_feature.shape[-1])
        return pooled_features

    @staticmethod
    def backward(ctx, grad_out):
        """
        :param grad_out: (N, out_x, out_y, out_z, C)
        :return grad_in: (npoints, C)
        """
        pts_idx_of_voxels, argmax, pool_method, num_pts, num_channels = ctx.roiaware_pool3d_for_backward
        grad_in = roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, num_pts, pool_method)
        return None, None, grad_in, None, None, None


def roiaware_pool3d_forward(rois, pts, pts_feature, out_size, max_pts_each_voxel=128, pool_method='max'):
    """The extension call of the reference (roiaware_pool3d_cuda.forward, roiaware_pool3d.cpp:25-62) with the three outputs
    returned instead of pre-allocated: pooled (N, ox, oy, oz, C) f32, argmax (same shape, int32; zeros for 'avg', which never
    reads it), pts_idx_of_voxels (N, ox, oy, oz, max_pts) int32."""
    assert rois.is_cuda and pts.is_cuda and pts_feature.is_cuda
    out_x, out_y, out_z = (out_size,) * 3 if isinstance(out_size, int) else out_size
    r, p, f = rois.contiguous().float(), pts.contiguous().float(), pts_feature.contiguous().float()
    n, m, c = r.shape[0], p.shape[0], f.shape[-1]
    method = {'max': 0, 'avg': 1}[pool_method]
    pooled = torch.empty((n, out_x, out_y, out_z, c), dtype=torch.float32, device=f.device)
    # allocated for both methods, as the reference does (roiaware_pool3d_utils.py:85); avg pooling leaves it zero-filled
    argmax = torch.empty((n, out_x, out_y, out_z, c), dtype=torch.int32, device=f.device)
    pts_idx = torch.empty((n, out_x, out_y, out_z, max_pts_each_voxel), dtype=torch.int32, device=f.device)
    L = _lib.lib()
    with torch.cuda.device(f.device):
        rc = L.lg_roiaware_pool3d_forward(_lib.ptr(r), n, _lib.ptr(p), m, _lib.ptr(f), c, out_x, out_y, out_z, max_pts_each_voxel,
                                          method, _lib.ptr(pooled), _lib.ptr(argmax), _lib.ptr(pts_idx), _lib.LG_FLAG_NONE,
                                          _lib.stream_ptr(f.device))
    _lib.check(rc, 'lg_roiaware_pool3d_forward')
    return pooled, argmax, pts_idx


def roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, num_pts, pool_method):
    """roiaware_pool3d_cuda.backward (roiaware_pool3d.cpp:64-95): -> grad_in (num_pts, C); pool_method 0 / 'max' or 1 / 'avg'."""
    method = {'max': 0, 'avg': 1}.get(pool_method, pool_method)
    g = grad_out.contiguous().float()
    n, out_x, out_y, out_z, max_pts = pts_idx_of_voxels.shape
    c = g.shape[-1]
    grad_in = g.new_zeros((num_pts, c))
    L = _lib.lib()
    with torch.cuda.device(g.device):
        rc = L.lg_roiaware_pool3d_backward(_lib.ptr(pts_idx_of_voxels), _lib.ptr(argmax), _lib.ptr(g), _lib.ptr(grad_in), n,
                                           out_x, out_y, out_z, c, max_pts, method, _lib.LG_FLAG_NONE, _lib.stream_ptr(g.device))
    _lib.check(rc, 'lg_roiaware_pool3d_backward')
    return grad_in
