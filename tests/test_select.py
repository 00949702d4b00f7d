"""lg_select_topk / lg_select_finish (the selection steps of model_nms_utils.py:6-25, batched) against the torch
operations they replace (`scores >= thresh` mask, torch.topk, gather; keep[:post], indices[keep], scores[selected]).
Bar: identical indices / counts / boxes / scores for distinct scores; for equal scores the documented rule (ascending
candidate index) and the same multiset of scores as torch.topk."""
import ctypes as C

import numpy as np
import pytest

from lidardetection_b200 import _lib


def test_select_c_abi_validation_without_gpu():
    L = _lib.lib()
    err = lambda: L.lg_last_error_string().decode()  # noqa: E731
    d = C.c_void_p(16)
    assert L.lg_select_workspace_bytes(64, 70400) == 64 * 70400 * 8 and L.lg_select_workspace_bytes(-1, 5) == 0
    assert L.lg_select_topk(d, -1, 10, 4, 0.0, 0, d, 70, 7, 1, d, d, d, d, 1 << 20, 0, None) == -1 and "invalid size" in err()
    assert L.lg_select_topk(d, 2, 10, 5000, 0.0, 0, d, 70, 7, 1, d, d, d, d, 1 << 20, 0, None) == -3 and "LG_SELECT_MAX_K" in err()
    assert L.lg_select_topk(None, 2, 10, 4, 0.0, 0, d, 70, 7, 1, d, d, d, d, 1 << 20, 0, None) == -1 and "null" in err()
    assert L.lg_select_topk(d, 2, 10, 4, 0.0, 0, d, 70, 7, 1, d, d, d, None, 0, 0, None) == -2
    assert L.lg_select_topk(None, 0, 10, 4, 0.0, 0, None, 70, 7, 1, None, None, None, None, 0, 0, None) == 0
    assert L.lg_select_finish(d, d, d, d, 2, 10, 4, -1, d, d, d, None) == -1
    assert L.lg_select_finish(None, None, None, None, 0, 10, 4, 3, None, None, None, None) == 0


gpu = pytest.mark.gpu


def native_topk(scores, boxes, k, thresh, ppf=1):
    import torch

    L = _lib.lib()
    P, N = scores.shape
    dev = scores.device
    top_idx = torch.empty((P, k), dtype=torch.int64, device=dev)
    counts = torch.empty((P,), dtype=torch.int32, device=dev)
    top_boxes = torch.empty((P, k, 7), dtype=torch.float32, device=dev)
    ws = torch.empty(L.lg_select_workspace_bytes(P, N), dtype=torch.uint8, device=dev)
    rc = L.lg_select_topk(_lib.ptr(scores), P, N, k, float(thresh if thresh is not None else 0.0), int(thresh is not None), _lib.ptr(boxes),
                          boxes.stride(0), boxes.stride(1), ppf, _lib.ptr(top_idx), _lib.ptr(counts), _lib.ptr(top_boxes), _lib.ptr(ws), ws.numel(), 0,
                          _lib.stream_ptr(dev))
    _lib.check(rc, "lg_select_topk")
    return top_idx, counts, top_boxes


@gpu
@pytest.mark.parametrize("P,N,k,thresh", [(3, 1000, 100, None), (5, 70400, 4096, 0.9), (2, 70400, 4096, None), (4, 5000, 4096, 0.3),
                                          (2, 37, 64, None), (1, 9000, 1, None), (3, 4096, 4096, 0.999999)])
def test_topk_matches_torch_for_distinct_scores(P, N, k, thresh):
    import torch

    g = torch.Generator().manual_seed(P * 1000 + N + k)
    scores = torch.stack([(torch.randperm(N, generator=g) + 1).float() / (N + 1) for _ in range(P)]).cuda()  # distinct by construction
    boxes = torch.randn((P, N, 9), generator=g).cuda()  # wider rows than 7: the strided slice callers pass
    kk = min(k, N)
    ti, cn, tb = native_topk(scores, boxes, kk, thresh)
    masked = scores if thresh is None else torch.where(scores >= thresh, scores, scores.new_full((), float("-inf")))
    ws, wi = torch.topk(masked, k=kk, dim=1)
    wc = (ws > float("-inf")).sum(1).to(torch.int32)
    assert torch.equal(cn, wc)
    for p in range(P):
        c = int(wc[p])
        assert torch.equal(ti[p, :c], wi[p, :c]) and bool((ti[p, c:] == 0).all())
        assert torch.equal(tb[p, :c], boxes[p, wi[p, :c], 0:7]) and bool((tb[p, c:] == 0).all())


@gpu
def test_topk_ties_special_values_and_shared_boxes():
    import torch

    # equal scores: ascending candidate index; the selected multiset equals torch's
    N, k = 20000, 4096
    scores = torch.full((2, N), 0.5).cuda()
    scores[1] = torch.randint(0, 7, (N,)).float().cuda() / 7  # 7 distinct values, thousands of ties at the cut
    boxes = torch.randn((1, N, 7)).cuda()
    ti, cn, tb = native_topk(scores, boxes, k, None, ppf=2)  # two "classes" share frame 0's boxes
    assert cn.tolist() == [k, k]
    assert torch.equal(ti[0], torch.arange(k, device="cuda"))
    ws, _ = torch.topk(scores[1], k)
    assert torch.equal(scores[1][ti[1]], ws)
    for v in scores[1][ti[1]].unique():  # inside a run of equal scores the indices ascend
        run = ti[1][scores[1][ti[1]] == v]
        assert bool((run[1:] > run[:-1]).all())
    cut = ws[-1]
    eq_all = (scores[1] == cut).nonzero().squeeze(1)
    eq_sel = ti[1][scores[1][ti[1]] == cut]
    assert torch.equal(eq_sel, eq_all[: len(eq_sel)])  # the lowest indices of the tied group are the ones taken
    assert torch.equal(tb[1], boxes[0, ti[1]])
    # special values: NaN never passes a threshold; -inf passes `>= -inf`; nothing passes -> counts 0
    s = torch.tensor([[0.3, float("nan"), float("inf"), -1.0, float("-inf"), 0.3, 2.0]]).cuda()
    b = torch.randn((1, 7, 7)).cuda()
    ti, cn, _ = native_topk(s, b, 7, -5.0)
    assert cn.item() == 5 and ti[0, :5].tolist() == [2, 6, 0, 5, 3]
    ti, cn, _ = native_topk(s, b, 7, None)
    assert cn.item() == 7 and ti[0].tolist() == [1, 2, 6, 0, 5, 3, 4]  # positive NaN first, like torch.topk
    ti, cn, tb = native_topk(s, b, 4, 100.0)
    assert cn.item() == 1 and ti[0].tolist() == [2, 0, 0, 0] and torch.equal(tb[0, 0], b[0, 2]) and bool((tb[0, 1:] == 0).all())  # +inf >= 100
    ti, cn, tb = native_topk(s[:, 3:6].contiguous(), b[:, 3:6].contiguous(), 3, 100.0)
    assert cn.item() == 0 and bool((ti == 0).all()) and bool((tb == 0).all())


@gpu
def test_front_end_native_equals_torch_path(monkeypatch):
    import torch

    from lidardetection_b200 import model_nms_utils as M, synth

    bx, sc = synth.cfg2(n_frames=6, n_boxes=1500, seed=77)
    boxes = torch.from_numpy(np.concatenate([bx, np.zeros((6, 1500, 2), np.float32)], 2)).cuda()
    scores = torch.from_numpy(sc).cuda()
    cfg = {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.05, "NMS_PRE_MAXSIZE": 1024, "NMS_POST_MAXSIZE": 40}
    a = M.class_agnostic_nms_batched(scores, boxes, cfg, score_thresh=0.3)
    monkeypatch.setattr(M, "LG_SELECT_MAX_K", 0)  # force the torch-ops path
    b = M.class_agnostic_nms_batched(scores, boxes, cfg, score_thresh=0.3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    cls = torch.rand((3, 1500, 4), generator=torch.Generator().manual_seed(5)).cuda()
    monkeypatch.setattr(M, "LG_SELECT_MAX_K", 4096)
    a = M.multi_classes_nms_batched(cls, boxes[:3], cfg, score_thresh=0.2)
    monkeypatch.setattr(M, "LG_SELECT_MAX_K", 0)
    b = M.multi_classes_nms_batched(cls, boxes[:3], cfg, score_thresh=0.2)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
