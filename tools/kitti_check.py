#!/usr/bin/env python
"""GPU box: our KITTI-eval kernels against (a) the oracle and (b) the reference numba kernel's cubin run live; bit level."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_kitti as MG  # noqa: E402
from lidardetection_b200.datasets.kitti.kitti_object_eval_python import eval as E, rotate_iou as R  # noqa: E402
from oracle import lg_oracle as O, ref_kitti  # noqa: E402


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def cmp(name, a, b, mask=None):
    ne = bits(a) != bits(b)
    ne &= ~(np.isnan(a) & np.isnan(b))
    if mask is not None:
        ne &= mask
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d = np.where(np.isfinite(d), d, 0)
    print(f"{name}: {int(ne.sum())} bit mismatches of {a.size}, max abs diff {d[ne].max() if ne.any() else 0:.3e}")
    return int(ne.sum())


for name, (b, q) in MG.cases().items():
    cnt = O.rotate_iou_eval_cnt(b, q)
    ok = cnt <= 8
    print(f"== {name}: {b.shape[0]} x {q.shape[0]}, pairs with > 8 polygon points (undefined in the reference): {int((~ok).sum())}")
    for c in MG.CRITERIA:
        ours = R.rotate_iou_gpu_eval(b, q, c)
        ref = ref_kitti.rotate_iou_gpu_eval(b, q, c)
        orc = O.rotate_iou_eval(b, q, c)
        cmp(f"  c={c:2d} ours vs oracle   ", ours, orc)
        cmp(f"  c={c:2d} ours vs reference", ours, ref, ok)
        cmp(f"  c={c:2d} oracle vs reference", orc, ref, ok)
G, D = MG.d3_case()
for c in MG.CRITERIA:
    cmp(f"d3 c={c:2d} ours vs oracle", E.d3_box_overlap(G, D, c), O.d3_box_overlap(G, D, c))
# parts
gts, dts = __import__("lidardetection_b200.synth", fromlist=["x"]).kitti_eval_frames(3769, 5)
ga = [{"name": [0] * len(g), "location": g[:, 0:3], "dimensions": g[:, 3:6], "rotation_y": g[:, 6]} for g in gts]
da = [{"name": [0] * len(d), "location": d[:, 0:3], "dimensions": d[:, 3:6], "rotation_y": d[:, 6]} for d in dts]
for metric in (1, 2):
    E.calculate_iou_partly(ga, da, metric)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ov, parts, ng, nd = E.calculate_iou_partly(ga, da, metric)
    dt_ = time.perf_counter() - t
    tot = sum(p.size for p in parts)
    print(f"calculate_iou_partly metric {metric}: 3769 frames, {len(parts)} parts, {tot} pairs, {dt_ * 1e3:.1f} ms host wall")
    p0 = E.d3_box_overlap(np.concatenate(gts[:75]), np.concatenate(dts[:75])) if metric == 2 else E.bev_box_overlap(
        np.concatenate(gts[:75])[:, MG.BEV_COLS], np.concatenate(dts[:75])[:, MG.BEV_COLS])
    cmp(f"  part 0 vs single call", parts[0].astype(np.float32), p0)
print("done")
