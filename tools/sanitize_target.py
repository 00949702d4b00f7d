#!/usr/bin/env python
"""A few small calls of every kernel, for `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` (ONE tool per gpurun call)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import model_nms_utils as MU, synth  # noqa: E402
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U  # noqa: E402
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU  # noqa: E402

cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
a, b = synth.clustered_pairs(150, 131, 1)
U.boxes_iou_bev(cu(a), cu(b)); U.boxes_iou3d_gpu(cu(a), cu(b[:20])); U.boxes_overlap_bev(cu(a[:3]), cu(b[:300].repeat(2, 0)[:257]))
a4, b4 = synth.cfg4(700)
U.boxes_iou3d_gpu(cu(a4), cu(b4)); U.boxes_iou_max(cu(a4), cu(b4), rows=True, cols=True)
r = np.random.default_rng(0)
near = a[:64] + r.normal(0, 1e-3, (64, 7)).astype(np.float32)
U.boxes_iou_bev(cu(a[:64]), cu(near))  # deferred (16-vertex / literal) paths
bx, sc = synth.nms_frames(3, 700, seed=2)
U.nms_gpu_batched(cu(bx), cu(sc), 0.1, counts=torch.tensor([700, 65, 0], dtype=torch.int32))
U.nms_gpu_batched(cu(bx), cu(sc), 0.1, full_mask=True)
U.nms_normal_gpu(cu(bx[0]), cu(sc[0]), 0.3)
MU.class_agnostic_nms_batched(cu(sc), cu(bx), {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.1, "NMS_PRE_MAXSIZE": 300, "NMS_POST_MAXSIZE": 50}, score_thresh=0.3)
pts, rois = synth.cfg3(n_frames=2, n_points=3001, n_rois=60, seed=5)
PU.points_in_boxes_gpu(cu(pts), cu(rois))
pts, rois = synth.cfg3(n_frames=1, n_points=4096, n_rois=300, seed=6)
PU.points_in_boxes_gpu(cu(pts), cu(rois))
PU.points_in_boxes_mask_gpu(cu(pts[0]), cu(rois[0]))
torch.cuda.synchronize()
print("sanitize target ok")
