"""Drop-in for pcdet/ops/roipoint_pool3d/roipoint_pool3d_utils.py (a "next" row): the PointRCNN
RoI point pooling layer, over lg_roipoint_pool3d_forward."""
import torch
import torch.nn as nn
from torch.autograd import Function

from ... import _lib


def enlarge_box3d(boxes3d, extra_width=(0, 0, 0)):
    """pcdet/utils/box_utils.py:136-149 (torch path): dims grow by extra_width, centre and heading unchanged."""
    large_boxes3d = boxes3d.clone()
    large_boxes3d[:, 3:6] += boxes3d.new_tensor(extra_width)[None, :]
    return large_boxes3d


class RoIPointPool3d(nn.Module):
    """roipoint_pool3d_utils.py:9-29."""

    def __init__(self, num_sampled_points=512, pool_extra_width=1.0):
        super().__init__()
        self.num_sampled_points = num_sampled_points
        self.pool_extra_width = pool_extra_width

    def forward(self, points, point_features, boxes3d):
        """
        Args:
            points: (B, N, 3)
            point_features: (B, N, C)
            boxes3d: (B, M, 7), [x, y, z, dx, dy, dz, heading]
        Returns:
            pooled_features: (B, M, 512, 3 + C)
            pooled_empty_flag: (B, M)
        """
        return RoIPointPool3dFunction.apply(points, point_features, boxes3d, self.pool_extra_width, self.num_sampled_points)


class RoIPointPool3dFunction(Function):
    """roipoint_pool3d_utils.py:32-66."""

    @staticmethod
    def forward(ctx, points, point_features, boxes3d, pool_extra_width, num_sampled_points=512):
        assert points.shape.__len__() == 3 and points.shape[2] == 3
        batch_size, boxes_num, feature_len = points.shape[0], boxes3d.shape[1], point_features.shape[2]
        pooled_boxes3d = enlarge_box3d(boxes3d.view(-1, 7), pool_extra_width).view(batch_size, -1, 7)
        return roipoint_pool3d_forward(points, pooled_boxes3d, point_features, num_sampled_points)

    @staticmethod
    def backward(ctx, grad_out):
        raise NotImplementedError


def roipoint_pool3d_forward(points, pooled_boxes3d, point_features, num_sampled_points=512):
    """The extension call of the reference (roipoint_pool3d_cuda.forward, roipoint_pool3d.cpp:24-58) on already enlarged
    boxes, outputs returned: pooled (B, M, S, 3 + C) f32, empty flag (B, M) int32."""
    assert points.is_cuda and pooled_boxes3d.is_cuda and point_features.is_cuda
    p, b, f = points.contiguous().float(), pooled_boxes3d.contiguous().float(), point_features.contiguous().float()
    batch_size, n, _ = p.shape
    m, c = b.shape[1], f.shape[2]
    pooled_features = torch.empty((batch_size, m, num_sampled_points, 3 + c), dtype=torch.float32, device=f.device)
    pooled_empty_flag = torch.empty((batch_size, m), dtype=torch.int32, device=f.device)
    L = _lib.lib()
    with torch.cuda.device(f.device):
        rc = L.lg_roipoint_pool3d_forward(_lib.ptr(p), _lib.ptr(b), _lib.ptr(f), batch_size, n, m, c, num_sampled_points,
                                          _lib.ptr(pooled_features), _lib.ptr(pooled_empty_flag), _lib.LG_FLAG_NONE,
                                          _lib.stream_ptr(f.device))
    _lib.check(rc, 'lg_roipoint_pool3d_forward')
    return pooled_features, pooled_empty_flag
