"""Drop-in for pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py (a "next" row).

The reference is a numba.cuda kernel (rotate_iou.py:260-291) behind a numpy-in / numpy-out wrapper that copies both box
sets to the device, launches 64-thread blocks and copies the matrix back (:293-330).  Here the same numbers come from
liblidargeom.so (lg_rotate_iou_eval: per-box records, exact-zero culling, queued polygon path); the arithmetic is that of
the reference kernel as built for sm_100a (include/lidargeom.h).  No numba, no CPU fallback.
"""
import numpy as np
import torch

from .... import _lib


def rotate_iou_eval_cuda(boxes, query_boxes, criterion=-1, flags=_lib.LG_FLAG_NONE):
    """device form: boxes (N, 5) / query_boxes (K, 5) float32 cuda tensors -> (N, K) float32 cuda tensor, stream-ordered"""
    assert boxes.is_cuda and query_boxes.is_cuda
    b, q = boxes.contiguous().float(), query_boxes.contiguous().float()
    assert b.dim() == 2 and b.shape[1] == 5 and q.dim() == 2 and q.shape[1] == 5
    n, k = b.shape[0], q.shape[0]
    out = torch.empty((n, k), dtype=torch.float32, device=b.device)
    if n == 0 or k == 0:
        return out
    L = _lib.lib()
    ws = torch.empty(L.lg_kitti_workspace_bytes(n, k, 0), dtype=torch.uint8, device=b.device)
    with torch.cuda.device(b.device):
        rc = L.lg_rotate_iou_eval(_lib.ptr(b), n, _lib.ptr(q), k, _lib.ptr(out), int(criterion), _lib.ptr(ws), ws.numel(), flags,
                                  _lib.stream_ptr(b.device))
    _lib.check(rc, "lg_rotate_iou_eval")
    return out


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """rotated box iou running in gpu (rotate_iou.py:293-330).

    Args:
        boxes (float array: [N, 5]): rbboxes. format: centers, dims, angles(clockwise when positive)
        query_boxes (float array: [K, 5])
        criterion: -1 iou, 0 inter / area(query box), 1 inter / area(box), else the intersection area
        device_id (int, optional): Defaults to 0.
    Returns:
        (N, K) float32 numpy array (the reference's `iou.astype(boxes.dtype)` sees the float32 copy of `boxes`)
    """
    boxes = np.asarray(boxes).astype(np.float32)
    query_boxes = np.asarray(query_boxes).astype(np.float32)
    N, K = boxes.shape[0], query_boxes.shape[0]
    if N == 0 or K == 0:
        return np.zeros((N, K), dtype=np.float32)
    dev = torch.device("cuda", device_id)
    out = rotate_iou_eval_cuda(torch.from_numpy(boxes.reshape(N, 5)).to(dev), torch.from_numpy(query_boxes.reshape(K, 5)).to(dev), criterion)
    return out.cpu().numpy()
