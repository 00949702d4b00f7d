"""Drop-in for pcdet/models/model_utils/model_nms_utils.py (the first "next" row): the
post-processing front end around NMS -- score threshold -> top-k -> NMS -> truncate -> map back.

`class_agnostic_nms` / `multi_classes_nms` keep the reference's names, signatures and return values
(model_nms_utils.py:6-25, 28-65).  The reference reaches them from a Python loop over the frames of a batch
(detector3d_template.py:190-260) and, for multi-head models, over heads and classes; every iteration issues a
chain of ~10 tiny launches and two host synchronisations.  `class_agnostic_nms_batched` /
`multi_classes_nms_batched` run the same selection for ALL frames (x classes) of a batch at once: one fused
threshold / sorted top-k / gather launch (lg_select_topk), one batched NMS call (lidargeom's lazy NMS) and one
truncate-and-map-back launch (lg_select_finish), no host synchronisation.  Equal scores are ordered by ascending
candidate index (torch.topk leaves that order unspecified).
"""
import torch

from .ops.iou3d_nms import iou3d_nms_utils


def _cfg(nms_config, key):
    return nms_config[key] if isinstance(nms_config, dict) else getattr(nms_config, key)


def _as_kwargs(nms_config):
    return dict(nms_config) if isinstance(nms_config, dict) else dict(vars(nms_config))


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """model_nms_utils.py:6-25 -- one frame.  Returns (selected indices into box_scores, their scores)."""
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = (box_scores >= score_thresh)
        box_scores = box_scores[scores_mask]
        box_preds = box_preds[scores_mask]

    selected = []
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(_cfg(nms_config, 'NMS_PRE_MAXSIZE'), box_scores.shape[0]))
        boxes_for_nms = box_preds[indices]
        keep_idx, selected_scores = getattr(iou3d_nms_utils, _cfg(nms_config, 'NMS_TYPE'))(
            boxes_for_nms[:, 0:7], box_scores_nms, _cfg(nms_config, 'NMS_THRESH'), **_as_kwargs(nms_config)
        )
        selected = indices[keep_idx[:_cfg(nms_config, 'NMS_POST_MAXSIZE')]]

    if score_thresh is not None:
        original_idxs = scores_mask.nonzero().view(-1)
        selected = original_idxs[selected]
    return selected, src_box_scores[selected]


def multi_classes_nms(cls_scores, box_preds, nms_config, score_thresh=None):
    """model_nms_utils.py:28-65 -- one frame, a Python loop over classes in the reference; here one batched call.

    Args:
        cls_scores: (N, num_class)
        box_preds: (N, 7 + C)
    Returns: pred_scores, pred_labels, pred_boxes (concatenated over classes, class-major, as the reference)
    (The reference leaves `cur_box_preds` unbound when score_thresh is None, model_nms_utils.py:41-51; here that
    case uses all boxes.)
    """
    sel, num, _ = multi_classes_nms_batched(cls_scores.unsqueeze(0), box_preds.unsqueeze(0), nms_config, score_thresh)
    num = num[0].tolist()
    pred_scores, pred_labels, pred_boxes = [], [], []
    for k, n in enumerate(num):
        idx = sel[0, k, :n]
        pred_scores.append(cls_scores[idx, k])
        pred_labels.append(cls_scores.new_ones(n).long() * k)
        pred_boxes.append(box_preds[idx])
    return torch.cat(pred_scores, dim=0), torch.cat(pred_labels, dim=0), torch.cat(pred_boxes, dim=0)


LG_SELECT_MAX_K = 4096  # lg_select_topk sorts the selection in shared memory


def _select_native(scores, box_preds, ppf, nms_config, score_thresh, k, post):
    """lg_select_topk -> batched NMS -> lg_select_finish: four launches of ours (records + lazy NMS in between), no torch op."""
    from . import _lib

    P, N = scores.shape
    dev = scores.device
    L = _lib.lib()
    sc = scores.contiguous()
    bp = box_preds if box_preds.stride(2) == 1 else box_preds.contiguous()
    top_idx = torch.empty((P, k), dtype=torch.int64, device=dev)
    counts = torch.empty((P,), dtype=torch.int32, device=dev)
    top_boxes = torch.empty((P, k, 7), dtype=torch.float32, device=dev)
    ws = torch.empty(L.lg_select_workspace_bytes(P, N), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        rc = L.lg_select_topk(_lib.ptr(sc), P, N, k, float(score_thresh if score_thresh is not None else 0.0), int(score_thresh is not None),
                              _lib.ptr(bp), bp.stride(0), bp.stride(1), ppf, _lib.ptr(top_idx), _lib.ptr(counts), _lib.ptr(top_boxes),
                              _lib.ptr(ws), ws.numel(), 0, st)
    _lib.check(rc, 'lg_select_topk')
    fn = {'nms_gpu': 'lg_nms_rotated_batched', 'nms_normal_gpu': 'lg_nms_normal_batched'}[_cfg(nms_config, 'NMS_TYPE')]
    # NMS_POST_MAXSIZE stops the NMS itself (model_nms_utils.py:20): keep keeps its (P, k) pitch, only its first `post` columns are written
    keep, num = iou3d_nms_utils._nms_call(fn, top_boxes, None, counts, float(_cfg(nms_config, 'NMS_THRESH')),
                                          buffers=iou3d_nms_utils._nms_buffers(fn, P, k, dev), max_keep=post)
    selected = torch.empty((P, post), dtype=torch.int64, device=dev)
    num_out = torch.empty((P,), dtype=torch.int32, device=dev)
    sel_scores = torch.empty((P, post), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.lg_select_finish(_lib.ptr(keep), _lib.ptr(num), _lib.ptr(top_idx), _lib.ptr(sc), P, N, k, post, _lib.ptr(selected),
                                _lib.ptr(num_out), _lib.ptr(sel_scores), st)
    _lib.check(rc, 'lg_select_finish')
    return selected, num_out, sel_scores


def _select_batched(scores, box_preds, ppf, nms_config, score_thresh):
    """scores (P, N); box_preds (P / ppf, N, 7 + C): problem p uses the boxes of frame p // ppf (ppf = classes per frame)
    -> selected (P, POST) int64 indices into N (-1 padded), num (P,) int32, scores of the selected"""
    P, N = scores.shape
    post = int(_cfg(nms_config, 'NMS_POST_MAXSIZE'))
    if N == 0 or P == 0:
        return (torch.full((P, post), -1, dtype=torch.int64, device=scores.device), torch.zeros(P, dtype=torch.int32, device=scores.device),
                scores.new_zeros((P, post)))
    k = min(int(_cfg(nms_config, 'NMS_PRE_MAXSIZE')), N)
    if (scores.is_cuda and scores.dtype == torch.float32 and box_preds.dtype == torch.float32 and k <= LG_SELECT_MAX_K and N < 2 ** 31
            and box_preds.shape[2] >= 7):
        return _select_native(scores, box_preds, ppf, nms_config, score_thresh, k, post)
    # larger k (or other dtypes): the same selection with torch ops around the batched NMS
    boxes7 = box_preds[:, :, 0:7]
    if ppf > 1:
        boxes7 = boxes7.unsqueeze(1).expand(P // ppf, ppf, N, 7).reshape(P, N, 7)
    masked = scores if score_thresh is None else torch.where(scores >= score_thresh, scores, scores.new_full((), float('-inf')))
    top_scores, top_idx = torch.topk(masked, k=k, dim=1)  # sorted, descending: already the order NMS needs
    counts = (top_scores > float('-inf')).sum(1).to(torch.int32)
    top_boxes = torch.gather(boxes7, 1, top_idx.unsqueeze(-1).expand(P, k, 7)).contiguous()
    fn = {'nms_gpu': 'lg_nms_rotated_batched', 'nms_normal_gpu': 'lg_nms_normal_batched'}[_cfg(nms_config, 'NMS_TYPE')]
    keep, num = iou3d_nms_utils._nms_call(fn, top_boxes.float(), None, counts.contiguous(), float(_cfg(nms_config, 'NMS_THRESH')))
    num = torch.clamp(num, max=post)
    keep = keep[:, :post]
    if keep.shape[1] < post:
        keep = torch.cat([keep, keep.new_full((P, post - keep.shape[1]), -1)], 1)
    valid = torch.arange(post, device=keep.device).unsqueeze(0) < num.unsqueeze(1)
    safe = torch.where(valid, keep, torch.zeros_like(keep))
    selected = torch.where(valid, torch.gather(top_idx, 1, safe), keep.new_full((), -1))
    sel_scores = torch.where(valid, torch.gather(scores, 1, torch.where(valid, selected, torch.zeros_like(selected))), scores.new_zeros(()))
    return selected, num, sel_scores


def class_agnostic_nms_batched(box_scores, box_preds, nms_config, score_thresh=None):
    """All frames of a batch at once.
    Args:
        box_scores: (B, N); box_preds: (B, N, 7 + C)
    Returns:
        selected (B, NMS_POST_MAXSIZE) int64 indices into N, in descending score order, -1 padded;
        num (B,) int32 number of selected boxes per frame; selected_scores (B, NMS_POST_MAXSIZE)
    Frame b's reference result (class_agnostic_nms) is selected[b, :num[b]].
    """
    assert box_scores.dim() == 2 and box_preds.dim() == 3 and box_preds.shape[:2] == box_scores.shape
    return _select_batched(box_scores, box_preds, 1, nms_config, score_thresh)


def multi_classes_nms_batched(cls_scores, box_preds, nms_config, score_thresh=None):
    """All frames x classes of a batch at once.
    Args:
        cls_scores: (B, N, num_class); box_preds: (B, N, 7 + C)
    Returns:
        selected (B, num_class, NMS_POST_MAXSIZE) int64 indices into N (-1 padded), num (B, num_class) int32, scores alike
    """
    B, N, C = cls_scores.shape
    scores = cls_scores.permute(0, 2, 1).reshape(B * C, N)
    sel, num, sc = _select_batched(scores, box_preds, C, nms_config, score_thresh)
    post = sel.shape[1]
    return sel.view(B, C, post), num.view(B, C), sc.view(B, C, post)
