"""GPU, full BASELINE.json sizes: size-independent properties where the oracle would take minutes."""
import numpy as np
import pytest
import torch

from lidardetection_b200 import sharded, synth
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU
from oracle import lg_oracle as O

pytestmark = pytest.mark.gpu


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_cfg2_full_batch_properties():
    """64 frames x 4096 boxes, thresh 0.01: batched == per-frame, NMS invariants hold, a frame is oracle-checked."""
    boxes, scores = synth.cfg2()
    tb, ts = cu(boxes), cu(scores)
    thr = 0.01
    keep, num = U.nms_gpu_batched(tb, ts, thr)
    assert keep.shape == (64, 4096)
    for f in range(64):
        n = int(num[f])
        k = keep[f, :n]
        assert n > 0 and bool((keep[f, n:] == -1).all())
        sc = ts[f][k]
        assert bool((sc[:-1] > sc[1:]).all())  # descending score order
        kb = tb[f][k]
        iou = U.boxes_iou_bev(kb, kb)
        assert bool((torch.triu(iou, 1) <= thr).all())  # survivors do not suppress each other
        if f % 8 == 0:
            # every removed box is suppressed by a kept box with a higher score
            removed = torch.ones(4096, dtype=torch.bool, device="cuda")
            removed[k] = False
            ridx = removed.nonzero().squeeze(1)
            cross = U.boxes_iou_bev(kb, tb[f][ridx])  # rows = kept (higher score first), cols = removed
            higher = sc.unsqueeze(1) > ts[f][ridx].unsqueeze(0)
            assert bool(((cross > thr) & higher).any(0).all())
            # idempotence
            k2 = U.nms_gpu(kb, sc, thr)[0]
            assert torch.equal(k2, torch.arange(n, device="cuda"))
            assert torch.equal(U.nms_gpu(tb[f], ts[f], thr)[0], k)
    for f in (0, 63):
        order = ts[f].sort(0, descending=True)[1].cpu().numpy()
        want = O.nms(boxes[f], scores[f], thr, order=order)
        assert np.array_equal(keep[f, : int(num[f])].cpu().numpy(), want)


def test_cfg5_multihead_batch():
    """NuScenes CBGS: (frames x classes) problems of 1000 boxes, thresh 0.2 (a 16-frame slice of the 256)."""
    b, s = synth.cfg5(n_frames=16)
    tb, ts = cu(b.reshape(-1, 1000, 7)), cu(s.reshape(-1, 1000))
    keep, num = U.nms_gpu_batched(tb, ts, 0.2)
    assert keep.shape == (160, 1000)
    for p in range(0, 160, 13):
        order = ts[p].sort(0, descending=True)[1].cpu().numpy()
        want = O.nms(b.reshape(-1, 1000, 7)[p], s.reshape(-1, 1000)[p], 0.2, order=order)
        assert np.array_equal(keep[p, : int(num[p])].cpu().numpy(), want)


def test_iou_self_and_range_properties():
    a, b = synth.cfg4(n=20000)
    ta, tb = cu(a), cu(b)
    iou = U.boxes_iou3d_gpu(ta, tb)
    bev = U.boxes_iou_bev(ta, tb)
    assert bool((iou >= 0).all()) and bool((iou <= 1.0001).all()) and bool(torch.isfinite(iou).all())
    assert bool(((iou > 0) <= (bev > 0)).all())  # 3D overlap implies BEV overlap
    frac = float((bev > 0).float().mean())
    assert 0.0002 < frac < 0.02
    # identical boxes -> 8 corner vertices -> IoU 1 up to the algorithm's own FP32 conditioning at +-75 m
    # (the reference returns the same not-quite-1 values: compare with the oracle, bound loosely in absolute terms)
    d = torch.diagonal(U.boxes_iou_bev(ta[:4096], ta[:4096]))
    assert bool(((d - 1).abs() <= 1e-4).all())
    want_d = np.array([O.iou_bev_pair(a[i], a[i], O.FLAVOR_CUDA) for i in range(0, 4096, 16)], np.float32)
    assert np.abs(d[::16].cpu().numpy() - want_d).max() <= 1e-5
    d3 = torch.diagonal(U.boxes_iou3d_gpu(ta[:4096], ta[:4096]))
    assert bool(((d3 - 1).abs() <= 1e-4).all())
    # row blocks are independent: any row sharding reproduces the unsharded matrix bit for bit
    blk, (s, e) = sharded.boxes_iou_sharded(ta, tb, kind="iou3d")
    assert (s, e) == (0, 20000) and torch.equal(blk, iou)
    parts = [U.boxes_iou3d_gpu(ta[i:j], tb) for i, j in ((0, 7001), (7001, 13000), (13000, 20000))]
    assert torch.equal(torch.cat(parts, 0), iou)
    # sampled rows against the oracle
    rows = np.arange(0, 20000, 997)
    want = O.boxes_iou3d(a[rows], b, O.FLAVOR_CUDA)
    assert np.abs(iou[rows].cpu().numpy() - want).max() <= 1e-5


def test_large_matrix_64bit_offsets():
    """> 2^31 output elements in ONE call: the reference's int32 index (kernel.cu:248,264) cannot do this."""
    free, _ = torch.cuda.mem_get_info()
    n, m = 66000, 33000  # 2.178e9 pairs, 8.7 GB
    if free < n * m * 4 + (2 << 30):
        pytest.skip("not enough free device memory")
    a, b = synth.cfg4(n=n)
    ta, tb = cu(a), cu(b[:m])
    iou = U.boxes_iou_bev(ta, tb)
    assert iou.numel() > 2 ** 31
    rows = np.array([0, 1, 32767, 65000, 65999])
    want = O.boxes_iou_bev(a[rows], b[:m], O.FLAVOR_CUDA)
    assert np.abs(iou[rows].cpu().numpy() - want).max() <= 1e-5
    tail = U.boxes_iou_bev(ta[65000:], tb)
    assert torch.equal(tail, iou[65000:])
    del iou


def test_points_idx_consistent_with_mask_form_and_batching():
    pts, rois = synth.cfg3(n_frames=8)
    tp, tr = cu(pts), cu(rois)
    idx = PU.points_in_boxes_gpu(tp, tr)
    for f in range(8):
        single = PU.points_in_boxes_gpu(tp[f:f + 1], tr[f:f + 1])
        assert torch.equal(single[0], idx[f])  # callers loop frames with B = 1 (point_head_template.py:78-89)
        mask = PU.points_in_boxes_mask_gpu(tp[f], tr[f], margin=1e-5)  # (T, M)
        anyhit = mask.max(0).values > 0
        first = torch.where(anyhit, mask.argmax(0).to(torch.int32), torch.full_like(idx[f], -1))
        assert torch.equal(first, idx[f])  # first set bit == lowest box index wins
    out = sharded.points_in_boxes_sharded(tp, tr)
    assert torch.equal(out, idx)
