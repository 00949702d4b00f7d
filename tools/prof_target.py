#!/usr/bin/env python
"""Small, short-running targets for `ncu --set full` (≈40 replays per launch): one op, a few launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import synth  # noqa: E402
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U  # noqa: E402
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "nms"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cu = lambda x: torch.from_numpy(x).cuda()
if op == "nms_full":
    b, s = synth.cfg2(16, 4096)
    tb, ts = cu(b), cu(s)
    for _ in range(iters):
        U.nms_gpu_batched(tb, ts, 0.01, full_mask=True)
elif op == "nms64":
    b, s = synth.cfg2(64, 4096)
    tb, ts = cu(b), cu(s)
    for _ in range(iters):
        U.nms_gpu_batched(tb, ts, 0.01)
elif op == "nms":
    b, s = synth.cfg2(16, 4096)
    tb, ts = cu(b), cu(s)
    for _ in range(iters):
        U.nms_gpu_batched(tb, ts, 0.01)
elif op == "iou_dense":
    a, b = synth.dense_overlap(4096, 4096)
    ta, tb = cu(a), cu(b)
    for _ in range(iters):
        U.boxes_iou_bev(ta, tb)
elif op == "iou_sparse":
    a, b = synth.cfg4(32768)
    ta, tb = cu(a), cu(b)
    for _ in range(iters):
        U.boxes_iou3d_gpu(ta, tb)
elif op == "iou_dense16k":  # bench shape of iou_dense
    a, b = synth.dense_overlap(16384, 16384)
    ta, tb = cu(a), cu(b)
    for _ in range(iters):
        U.boxes_iou_bev(ta, tb)
elif op == "iou_cfg4":  # bench shape: one 25,000-row shard x 200,000 boxes, fused 3-D IoU (20 GB of output)
    a, b = synth.cfg4(200_000)
    ta, tb = cu(a[:25000]), cu(b)
    for _ in range(iters):
        out = U.boxes_iou3d_gpu(ta, tb)
        del out
elif op == "iou_cfg1":
    a, b = synth.cfg1()
    ta, tb = cu(a), cu(b)
    for _ in range(iters):
        U.boxes_iou_bev(ta, tb)
elif op == "nms_cfg5":
    b, s = synth.cfg5(256, 10, 1000)
    tb, ts = cu(b.reshape(-1, 1000, 7)), cu(s.reshape(-1, 1000))
    for _ in range(iters):
        U.nms_gpu_batched(tb, ts, 0.2)
elif op == "pib":
    p, r = synth.cfg3(256)
    tp, tr = cu(p), cu(r)
    for _ in range(iters):
        PU.points_in_boxes_gpu(tp, tr)
elif op == "pib4096":  # the bench's shape: one CTA per frame
    import numpy as np

    p, r = synth.cfg3(64)
    tp, tr = cu(np.tile(p, (64, 1, 1))), cu(np.tile(r, (64, 1, 1)))
    for _ in range(iters):
        PU.points_in_boxes_gpu(tp, tr)
elif op == "kitti":
    import numpy as np

    from lidardetection_b200.datasets.kitti.kitti_object_eval_python import eval as E

    gts, dts = synth.kitti_eval_frames(3769, 5000)
    parts = E.get_split_parts(3769, 50)
    gc, dc, i = [], [], 0
    for n in parts:
        gc.append(sum(len(x) for x in gts[i:i + n]))
        dc.append(sum(len(x) for x in dts[i:i + n]))
        i += n
    g, d = cu(np.concatenate(gts)), cu(np.concatenate(dts))
    for _ in range(iters):
        E.kitti_overlaps_parts_cuda(g, d, gc, dc, 2)
elif op in ("roiaware", "roipoint"):
    from lidardetection_b200.ops.roipoint_pool3d import roipoint_pool3d_utils as RU

    pts, rois, feat = synth.pool_case(16384, 128, 128)
    tp, tr, tf = cu(pts), cu(rois), cu(feat)
    for _ in range(iters):
        if op == "roiaware":
            po, am, pi = PU.roiaware_pool3d_forward(tr, tp, tf, 12, 128, "max")
            PU.roiaware_pool3d_backward(pi, am, po, 16384, "max")
        else:
            RU.roipoint_pool3d_forward(tp[None], tr[None], tf[None], 512)
torch.cuda.synchronize()
print("ok", op)
