"""ctypes shim with the interface of the reference's compiled extension `iou3d_nms_cuda`
(pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17): same function names, positional arguments, in-place outputs and
return values, over the C ABI of liblidargeom.so (include/lidargeom.h).  See lidardetection_b200/compat/__init__.py.
"""
import torch

from .. import _lib


def _check_boxes(t, name):
    # iou3d_nms.cpp:14-38 (CHECK_INPUT): CUDA, contiguous; here also float32 (the extension reads `float*`)
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor")


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _iou(fn, boxes_a, boxes_b, ans, flags=_lib.LG_FLAG_NONE):
    n, m = boxes_a.shape[0], boxes_b.shape[0]
    if n == 0 or m == 0:
        return 1
    L = _lib.lib()
    with torch.cuda.device(boxes_a.device):
        ws = _ws(L.lg_iou_workspace_bytes(n, m), boxes_a.device)
        rc = getattr(L, fn)(_lib.ptr(boxes_a), n, _lib.ptr(boxes_b), m, _lib.ptr(ans), ans.stride(0), _lib.ptr(ws), ws.numel(), flags,
                            _lib.stream_ptr(boxes_a.device))
    _lib.check(rc, fn)
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d_nms.cpp:49-68: boxes_a (N, 7), boxes_b (M, 7) -> ans_overlap (N, M) written in place; returns 1."""
    _check_boxes(boxes_a, "boxes_a"), _check_boxes(boxes_b, "boxes_b"), _check_boxes(ans_overlap, "ans_overlap")
    return _iou("lg_boxes_overlap_bev", boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """iou3d_nms.cpp:70-88: -> ans_iou (N, M) written in place; returns 1."""
    _check_boxes(boxes_a, "boxes_a"), _check_boxes(boxes_b, "boxes_b"), _check_boxes(ans_iou, "ans_iou")
    return _iou("lg_boxes_iou_bev", boxes_a, boxes_b, ans_iou)


def _nms(fn, boxes, keep, thresh):
    _check_boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise ValueError("keep must be a contiguous CPU LongTensor (iou3d_nms.cpp:94)")
    n = boxes.shape[0]
    if n == 0:
        return 0
    L = _lib.lib()
    dev = boxes.device
    with torch.cuda.device(dev):
        ws = _ws(L.lg_nms_workspace_bytes_ex(1, n, 1 if fn == "lg_nms_normal" else 0, 0), dev)
        keep_dev = torch.empty(n, dtype=torch.int64, device=dev)
        num = torch.zeros(1, dtype=torch.int32, device=dev)
        rc = getattr(L, fn)(_lib.ptr(boxes), None, n, float(thresh), _lib.ptr(ws), ws.numel(), _lib.ptr(keep_dev), _lib.ptr(num), 0,
                            _lib.stream_ptr(dev))
    _lib.check(rc, fn)
    k = int(num.item())  # the extension's signature forces this round trip: keep lives on the host, the count is returned by value
    keep[:k] = keep_dev[:k].cpu()
    return k


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:90-136: boxes (N, 7) CUDA, sorted by descending score; keep LongTensor(N) on the HOST receives the kept
    positions; returns their count."""
    return _nms("lg_nms_rotated", boxes, keep, nms_overlap_thresh)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:139-186 (axis-aligned BEV IoU)."""
    return _nms("lg_nms_normal", boxes, keep, nms_overlap_thresh)


def boxes_iou_bev_cpu(boxes_a_tensor, boxes_b_tensor, ans_iou_tensor):
    """iou3d_cpu.cpp:232-252: CPU float tensors (N, 7), (M, 7) -> ans_iou (N, M) written in place; returns 1.
    Computed on the GPU with the CPU build's arithmetic (LG_FLAG_STRICT_FP32: un-contracted FP32, glibc sinf / cosf)."""
    for t in (boxes_a_tensor, boxes_b_tensor, ans_iou_tensor):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("boxes_iou_bev_cpu takes contiguous float32 CPU tensors (iou3d_cpu.cpp:238-240)")
    _lib.require_usable_cuda("boxes_iou_bev_cpu")
    dev = torch.device("cuda", torch.cuda.current_device())
    a, b = boxes_a_tensor.to(dev), boxes_b_tensor.to(dev)
    out = torch.empty(ans_iou_tensor.shape, dtype=torch.float32, device=dev)
    _iou("lg_boxes_iou_bev", a, b, out, _lib.LG_FLAG_STRICT_FP32)
    ans_iou_tensor.copy_(out)
    return 1
