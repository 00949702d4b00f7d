#!/usr/bin/env python
"""Developer tool (one GPU): randomised differential checks of the paths that have two formulations --
lazy NMS (every launch variant: clusters, 512 / 256 threads) against the mask + sweep formulation, and the two-phase IoU sweep (also
with a list that overflows) against the one-kernel sweep.  Bit-exact equality is the bar.   python tools/fuzz_gpu.py [seconds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import _lib, synth  # noqa: E402
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(12345)
cu = lambda x: torch.from_numpy(x).cuda()  # noqa: E731
t_end = time.time() + budget
n_nms = n_iou = 0
while time.time() < t_end:
    # ---- NMS: lazy == full mask, with ragged counts and NMS_POST_MAXSIZE
    P = int(rng.choice([1, 2, 5, 17, 64, 150, 300, 700]))
    N = int(rng.choice([1, 31, 33, 100, 513, 1000, 1536, 2049, 4096, 6000]))
    if P * N > 1_500_000:
        P = max(1, 1_500_000 // N)
    thr = float(rng.choice([0.01, 0.1, 0.2, 0.31, 0.5, 0.7]))
    seed = int(rng.integers(1 << 30))
    boxes, scores = synth.nms_frames(P, N, seed=seed, k_range=(3, 40))
    if rng.random() < 0.3:  # exact duplicates and ties
        boxes[:, N // 2:] = boxes[:, : N - N // 2]
        scores[:, N // 3:] = scores[:, : N - N // 3]
    counts = torch.from_numpy(rng.integers(0, N + 1, size=P).astype(np.int32)) if rng.random() < 0.5 else None
    mk = int(rng.choice([0, 1, 7, 83, 500])) if rng.random() < 0.5 else None
    if mk == 0:
        mk = None
    tb, ts = cu(boxes), cu(scores)
    k0, n0 = U.nms_gpu_batched(tb, ts, thr, counts, full_mask=True, max_keep=mk)
    k1, n1 = U.nms_gpu_batched(tb, ts, thr, counts, max_keep=mk)
    assert torch.equal(n0, n1), ("nms counts", P, N, thr, seed, mk)
    assert torch.equal(k0, k1), ("nms keep", P, N, thr, seed, mk)
    n_nms += 1
    # ---- IoU: two-phase == one kernel == overflowing list
    if n_nms % 4 == 0:
        n = int(rng.choice([8200, 9001, 12000, 16385]))
        m = int(rng.choice([8200, 8193, 9000, 30011]))
        if n * m < (1 << 26):
            m = (1 << 26) // n + 77
        a, b = synth.cfg4(max(n, m), seed=seed % 1000)
        if rng.random() < 0.3:
            a, _ = synth.clustered_pairs(max(n, m), 8, seed=seed % 1000)  # moderately dense rows against sparse columns
        ta, tb2 = cu(a[:n]), cu(b[:m])
        fn = str(rng.choice(["lg_boxes_iou3d", "lg_boxes_iou_bev", "lg_boxes_overlap_bev"]))
        one = U._iou_call(fn, ta, tb2, flags=_lib.LG_FLAG_IOU_ONE_KERNEL)
        two = U._iou_call(fn, ta, tb2)
        assert torch.equal(one, two), ("iou two-phase", fn, n, m, seed)
        small = U._iou_call(fn, ta, tb2, flags=_lib.LG_FLAG_IOU_SMALL_LIST)
        assert torch.equal(one, small), ("iou small list", fn, n, m, seed)
        del one, two, small
        n_iou += 1
print(f"fuzz ok: {n_nms} NMS batches (lazy == full mask), {n_iou} IoU matrices (two-phase == small list == one kernel)")
