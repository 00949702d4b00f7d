#!/usr/bin/env python
"""Developer tool: one-kernel sweep (LG_FLAG_IOU_ONE_KERNEL) against the two-phase sweep over matrix sizes, sparse (cfg4-like) inputs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import _lib, synth  # noqa: E402
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (4096, 5793, 8192, 16384, 32768):
    a, b = synth.cfg4(n, seed=n)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty((n, n), dtype=torch.float32, device="cuda")
    res = []
    for flags in (_lib.LG_FLAG_IOU_ONE_KERNEL, _lib.LG_FLAG_NONE):
        for _ in range(3):
            U._iou_call("lg_boxes_iou3d", ta, tb, flags=flags, out=out)
        ms = 0.0
        for _ in range(10):
            flush.zero_()
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            U._iou_call("lg_boxes_iou3d", ta, tb, flags=flags, out=out)
            e.record()
            e.synchronize()
            ms += s.elapsed_time(e)
        res.append(ms / 10)
    print(f"n = m = {n:6d}  pairs {n * n:.3g}  one-kernel {res[0]:8.4f} ms   two-phase {res[1]:8.4f} ms   ratio {res[0] / res[1]:.3f}", flush=True)
