"""ctypes shim with the interface of the reference's compiled extension `roiaware_pool3d_cuda`
(pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177): forward, backward, points_in_boxes_gpu, points_in_boxes_cpu --
same positional arguments, in-place outputs and return values, over the C ABI of liblidargeom.so.
"""
import torch

from .. import _lib


def _cuda_contig(t, name, dtype):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise ValueError(f"{name} must be a contiguous {dtype} CUDA tensor")


def points_in_boxes_gpu(boxes_tensor, pts_tensor, box_idx_of_points_tensor):
    """roiaware_pool3d.cpp:98-118: boxes (B, T, 7), pts (B, M, 3) -> box_idx_of_points (B, M) int32 written in place
    (first containing box or -1); returns 1."""
    _cuda_contig(boxes_tensor, "boxes", torch.float32), _cuda_contig(pts_tensor, "pts", torch.float32)
    _cuda_contig(box_idx_of_points_tensor, "box_idx_of_points", torch.int32)
    B, T, M = boxes_tensor.shape[0], boxes_tensor.shape[1], pts_tensor.shape[1]
    if B == 0 or M == 0:
        return 1
    L = _lib.lib()
    with torch.cuda.device(pts_tensor.device):
        rc = L.lg_points_in_boxes(_lib.ptr(boxes_tensor), _lib.ptr(pts_tensor), _lib.ptr(box_idx_of_points_tensor), B, T, M, None, 0,
                                  _lib.LG_FLAG_NONE, _lib.stream_ptr(pts_tensor.device))
    _lib.check(rc, "lg_points_in_boxes")
    return 1


def points_in_boxes_cpu(boxes_tensor, pts_tensor, pts_indices_tensor):
    """roiaware_pool3d.cpp:121-168: CPU tensors boxes (N, 7), pts (M, 3) -> pts_indices (N, M) int32 0/1 written in place
    (MARGIN 1e-2); returns 1.  Computed on the GPU with the CPU build's arithmetic."""
    for t in (boxes_tensor, pts_tensor, pts_indices_tensor):
        if t.is_cuda or not t.is_contiguous():
            raise ValueError("points_in_boxes_cpu takes contiguous CPU tensors (roiaware_pool3d.cpp:127-129)")
    n, m = boxes_tensor.shape[0], pts_tensor.shape[0]
    if n == 0 or m == 0:
        return 1
    _lib.require_usable_cuda("points_in_boxes_cpu")
    dev = torch.device("cuda", torch.cuda.current_device())
    b, p = boxes_tensor.float().to(dev), pts_tensor.float().to(dev)
    out = torch.empty((n, m), dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.lg_points_in_boxes_mask(_lib.ptr(b), n, _lib.ptr(p), m, _lib.ptr(out), 1e-2, _lib.LG_FLAG_STRICT_FP32, _lib.stream_ptr(dev))
    _lib.check(rc, "lg_points_in_boxes_mask")
    pts_indices_tensor.copy_(out)
    return 1


def forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled_features, pool_method):
    """roiaware_pool3d.cpp:25-62: rois (N, 7), pts (M, 3), pts_feature (M, C); outputs written in place: argmax and
    pooled_features (N, ox, oy, oz, C), pts_idx_of_voxels (N, ox, oy, oz, max_pts); pool_method 0 = max, 1 = avg."""
    for t, name in ((rois, "rois"), (pts, "pts"), (pts_feature, "pts_feature"), (pooled_features, "pooled_features")):
        _cuda_contig(t, name, torch.float32)
    _cuda_contig(argmax, "argmax", torch.int32), _cuda_contig(pts_idx_of_voxels, "pts_idx_of_voxels", torch.int32)
    n, m, c = rois.shape[0], pts.shape[0], pts_feature.shape[1]
    _, ox, oy, oz, max_pts = pts_idx_of_voxels.shape
    if n == 0:
        return 1
    L = _lib.lib()
    with torch.cuda.device(pts.device):
        rc = L.lg_roiaware_pool3d_forward(_lib.ptr(rois), n, _lib.ptr(pts), m, _lib.ptr(pts_feature), c, ox, oy, oz, max_pts, int(pool_method),
                                          _lib.ptr(pooled_features), _lib.ptr(argmax), _lib.ptr(pts_idx_of_voxels), _lib.LG_FLAG_NONE,
                                          _lib.stream_ptr(pts.device))
    _lib.check(rc, "lg_roiaware_pool3d_forward")
    return 1


def backward(pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method):
    """roiaware_pool3d.cpp:64-95: accumulates into grad_in (num_pts, C), zero-filled by the caller."""
    _cuda_contig(pts_idx_of_voxels, "pts_idx_of_voxels", torch.int32), _cuda_contig(argmax, "argmax", torch.int32)
    _cuda_contig(grad_out, "grad_out", torch.float32), _cuda_contig(grad_in, "grad_in", torch.float32)
    n, ox, oy, oz, max_pts = pts_idx_of_voxels.shape
    c = grad_out.shape[4]
    if n == 0:
        return 1
    L = _lib.lib()
    with torch.cuda.device(grad_out.device):
        rc = L.lg_roiaware_pool3d_backward(_lib.ptr(pts_idx_of_voxels), _lib.ptr(argmax), _lib.ptr(grad_out), _lib.ptr(grad_in), n, ox, oy, oz,
                                           c, max_pts, int(pool_method), _lib.LG_FLAG_NONE, _lib.stream_ptr(grad_out.device))
    _lib.check(rc, "lg_roiaware_pool3d_backward")
    return 1
