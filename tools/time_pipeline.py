#!/usr/bin/env python
"""Developer tool: frames/s of HostNmsPipeline (pinned host -> keep lists in pinned host memory) on nms_cfg2 for 1..4 batches in flight."""
import sys, time, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidardetection_b200 import synth
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
b, s = synth.cfg2(64, 4096)
hb, hs = torch.from_numpy(b).pin_memory(), torch.from_numpy(s).pin_memory()
for depth in (1, 2, 3, 4):
    pipe = U.HostNmsPipeline(64, 4096, 0.01, max_keep=500, depth=depth)
    for _ in range(5):
        pipe.result(pipe.submit(hb, hs))
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        infl = []
        t0 = time.perf_counter()
        for _ in range(50):
            if len(infl) == depth:
                pipe.result(infl.pop(0))
            infl.append(pipe.submit(hb, hs))
        for t in infl:
            pipe.result(t)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 50)
    print(f"depth {depth}: {best * 1e3:.4f} ms/step  {64 / best:.0f} frames/s")
