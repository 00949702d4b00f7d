"""Drop-in for the overlap functions of pcdet/datasets/kitti/kitti_object_eval_python/eval.py:
image_box_overlap (:80-108), bev_box_overlap (:111-113), d3_box_overlap (:116-155) and calculate_iou_partly (:340-414).

The reference runs, for each of ~50 parts, H2D -> numba.cuda kernel -> D2H (-> a numba CPU pass over the matrix for 3-D).
`calculate_iou_partly` here concatenates the annotations once and evaluates ALL parts in one launch
(lg_kitti_overlaps_parts); the statistics half of eval.py (compute_statistics_jit, eval_class, ...) consumes the returned
per-image matrices unchanged and is out of scope.
"""
import threading

import numpy as np
import torch

from .... import _lib
from .rotate_iou import rotate_iou_gpu_eval


def get_split_parts(num, num_part):
    """eval.py:278-287"""
    same_part = num // num_part
    remain_num = num % num_part
    if same_part == 0:
        return [num]
    if remain_num == 0:
        return [same_part] * num_part
    return [same_part] * num_part + [remain_num]


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """eval.py:80-108 (numba CPU in the reference; axis-aligned image boxes, not part of the rotated path): vectorised numpy
    with the reference's operation order, dtype of `boxes`."""
    boxes = np.asarray(boxes)
    query_boxes = np.asarray(query_boxes)
    N, K = boxes.shape[0], query_boxes.shape[0]
    overlaps = np.zeros((N, K), dtype=boxes.dtype)
    if N == 0 or K == 0:
        return overlaps
    q = query_boxes.astype(np.result_type(boxes.dtype, query_boxes.dtype), copy=False)
    qbox_area = (q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1])
    iw = np.minimum(boxes[:, None, 2], q[None, :, 2]) - np.maximum(boxes[:, None, 0], q[None, :, 0])
    ih = np.minimum(boxes[:, None, 3], q[None, :, 3]) - np.maximum(boxes[:, None, 1], q[None, :, 1])
    box_area = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))[:, None]
    inter = iw * ih
    if criterion == -1:
        ua = box_area + qbox_area[None, :] - inter
    elif criterion == 0:
        ua = np.broadcast_to(box_area, inter.shape)
    elif criterion == 1:
        ua = np.broadcast_to(qbox_area[None, :], inter.shape)
    else:
        ua = np.ones_like(inter)
    ok = (iw > 0) & (ih > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        overlaps[ok] = (inter[ok] / ua[ok]).astype(boxes.dtype)
    return overlaps


def bev_box_overlap(boxes, qboxes, criterion=-1):
    """eval.py:111-113"""
    return rotate_iou_gpu_eval(boxes, qboxes, criterion)


def d3_box_overlap(boxes, qboxes, criterion=-1, device_id=0):
    """eval.py:150-155: float64 CAMERA boxes (x, y, z, l, h, w, ry) -> (N, K) float32.  One fused launch (lg_d3_box_overlap)
    instead of rotate_iou_gpu_eval(criterion=2) + the numba CPU pass d3_box_overlap_kernel."""
    boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
    qboxes = np.ascontiguousarray(np.asarray(qboxes, dtype=np.float64))
    N, K = boxes.shape[0], qboxes.shape[0]
    if N == 0 or K == 0:
        return np.zeros((N, K), dtype=np.float32)
    dev = torch.device("cuda", device_id)
    return d3_box_overlap_cuda(torch.from_numpy(boxes.reshape(N, 7)).to(dev), torch.from_numpy(qboxes.reshape(K, 7)).to(dev), criterion).cpu().numpy()


def d3_box_overlap_cuda(boxes, qboxes, criterion=-1, flags=_lib.LG_FLAG_NONE):
    """device form: (N, 7) / (K, 7) float64 cuda tensors -> (N, K) float32 cuda tensor"""
    assert boxes.is_cuda and qboxes.is_cuda and boxes.dtype == torch.float64 and qboxes.dtype == torch.float64
    b, q = boxes.contiguous(), qboxes.contiguous()
    n, k = b.shape[0], q.shape[0]
    out = torch.empty((n, k), dtype=torch.float32, device=b.device)
    if n == 0 or k == 0:
        return out
    L = _lib.lib()
    ws = torch.empty(L.lg_kitti_workspace_bytes(n, k, 0), dtype=torch.uint8, device=b.device)
    with torch.cuda.device(b.device):
        rc = L.lg_d3_box_overlap(_lib.ptr(b), n, _lib.ptr(q), k, _lib.ptr(out), int(criterion), _lib.ptr(ws), ws.numel(), flags,
                                 _lib.stream_ptr(b.device))
    _lib.check(rc, "lg_d3_box_overlap")
    return out


def kitti_overlaps_parts_cuda(gt_boxes, dt_boxes, gt_counts, dt_counts, metric, criterion=-1, flags=_lib.LG_FLAG_NONE):
    """All parts in one launch.  gt_boxes (sum G_p, 7) / dt_boxes (sum D_p, 7) float64 cuda, gt_counts / dt_counts: per-part row
    counts (host sequences).  Returns (out, out_off): the concatenated row-major (G_p x D_p) float32 matrices on the device and
    their int64 offsets on the host (len P + 1)."""
    assert gt_boxes.is_cuda and dt_boxes.is_cuda and gt_boxes.dtype == torch.float64 and dt_boxes.dtype == torch.float64
    g, d = gt_boxes.contiguous(), dt_boxes.contiguous()
    gc = np.asarray(gt_counts, dtype=np.int64)
    dc = np.asarray(dt_counts, dtype=np.int64)
    assert gc.shape == dc.shape and gc.sum() == g.shape[0] and dc.sum() == d.shape[0]
    P = len(gc)
    offs = np.zeros((3, P + 1), dtype=np.int64)
    offs[0, 1:] = np.cumsum(gc)
    offs[1, 1:] = np.cumsum(dc)
    offs[2, 1:] = np.cumsum(gc * dc)
    total = int(offs[2, -1])
    out = torch.empty(total, dtype=torch.float32, device=g.device)
    if total == 0:
        return out, offs[2]
    L = _lib.lib()
    doffs = torch.from_numpy(offs).to(g.device, non_blocking=True)
    ws = torch.empty(L.lg_kitti_workspace_bytes(g.shape[0], d.shape[0], P), dtype=torch.uint8, device=g.device)
    with torch.cuda.device(g.device):
        rc = L.lg_kitti_overlaps_parts(_lib.ptr(g), g.shape[0], _lib.ptr(d), d.shape[0], _lib.ptr(doffs[0]), _lib.ptr(doffs[1]),
                                       _lib.ptr(doffs[2]), P, total, int(metric), int(criterion), _lib.ptr(out), _lib.ptr(ws),
                                       ws.numel(), flags, _lib.stream_ptr(g.device))
    _lib.check(rc, "lg_kitti_overlaps_parts")
    return out, offs[2]


_tls = threading.local()


def _to_host_float64(t):
    """1-D float32 cuda tensor -> float64 numpy array.  D2H goes through a pinned staging buffer kept per thread (pageable
    D2H of ~100 MB costs more than the kernel by two orders of magnitude); the float64 copy the reference hands out
    (`.astype(np.float64)`, eval.py:379,392) is made by torch's multi-threaded CPU cast into fresh memory."""
    n = t.numel()
    buf = getattr(_tls, "pinned", None)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 20), dtype=torch.float32, pin_memory=True)
        _tls.pinned = buf
    buf[:n].copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf[:n].to(torch.float64).numpy()


def _camera_boxes(annos):
    if len(annos) == 0:
        return np.zeros((0, 7), dtype=np.float64)
    loc = np.concatenate([a["location"] for a in annos], 0)
    dims = np.concatenate([a["dimensions"] for a in annos], 0)
    rots = np.concatenate([a["rotation_y"] for a in annos], 0)
    return np.concatenate([loc, dims, rots[..., np.newaxis]], axis=1).astype(np.float64).reshape(-1, 7)


def calculate_iou_partly(gt_annos, dt_annos, metric, num_parts=50, device_id=0):
    """fast iou algorithm (eval.py:340-414).  Same arguments and return value as the reference:
    (overlaps, parted_overlaps, total_gt_num, total_dt_num) with float64 matrices; metric 0: bbox, 1: bev, 2: 3d."""
    assert len(gt_annos) == len(dt_annos)
    total_dt_num = np.stack([len(a["name"]) for a in dt_annos], 0)
    total_gt_num = np.stack([len(a["name"]) for a in gt_annos], 0)
    num_examples = len(gt_annos)
    split_parts = get_split_parts(num_examples, num_parts)
    parted_overlaps = []
    if metric == 0:
        example_idx = 0
        for num_part in split_parts:
            gt_boxes = np.concatenate([a["bbox"] for a in gt_annos[example_idx:example_idx + num_part]], 0)
            dt_boxes = np.concatenate([a["bbox"] for a in dt_annos[example_idx:example_idx + num_part]], 0)
            parted_overlaps.append(image_box_overlap(gt_boxes, dt_boxes))
            example_idx += num_part
    elif metric in (1, 2):
        gc, dc, example_idx = [], [], 0
        for num_part in split_parts:
            gc.append(int(total_gt_num[example_idx:example_idx + num_part].sum()))
            dc.append(int(total_dt_num[example_idx:example_idx + num_part].sum()))
            example_idx += num_part
        dev = torch.device("cuda", device_id)
        g = torch.from_numpy(_camera_boxes(gt_annos)).to(dev)
        d = torch.from_numpy(_camera_boxes(dt_annos)).to(dev)
        out, off = kitti_overlaps_parts_cuda(g, d, gc, dc, metric)
        # > 99 % of a KITTI part's entries are exactly +0.0: bring back only the others (selected by bit pattern, so -0.0 and NaN
        # survive) and scatter them into calloc'ed float64 memory instead of moving and converting the whole matrix
        nz = torch.nonzero(out.view(torch.int32)).squeeze(1)
        if nz.numel() * 8 <= out.numel():
            flat = np.zeros(out.numel(), dtype=np.float64)
            flat[nz.cpu().numpy()] = out[nz].cpu().numpy()
        else:
            flat = _to_host_float64(out)
        for p in range(len(split_parts)):
            parted_overlaps.append(flat[off[p]:off[p + 1]].reshape(gc[p], dc[p]))
    else:
        raise ValueError("unknown metric")
    overlaps = []
    example_idx = 0
    for j, num_part in enumerate(split_parts):
        gt_num_idx, dt_num_idx = 0, 0
        for i in range(num_part):
            gt_box_num = total_gt_num[example_idx + i]
            dt_box_num = total_dt_num[example_idx + i]
            overlaps.append(parted_overlaps[j][gt_num_idx:gt_num_idx + gt_box_num, dt_num_idx:dt_num_idx + dt_box_num])
            gt_num_idx += gt_box_num
            dt_num_idx += dt_box_num
        example_idx += num_part
    return overlaps, parted_overlaps, total_gt_num, total_dt_num
