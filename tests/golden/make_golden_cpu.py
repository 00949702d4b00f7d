#!/usr/bin/env python
"""Generate tests/golden/golden_cpu.npz from the UNMODIFIED reference CPU code (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the dev container (no GPU needed):

    python oracle/build_ref.py && python tests/golden/make_golden_cpu.py

Contents: inputs and the reference's outputs of boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252) and
points_in_boxes_cpu (roiaware_pool3d.cpp:143-168) on seeded synthetic boxes incl. hand-made edge cases.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidardetection_b200 import synth  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

PI = np.pi
# hand-made cases against a = [0,0,0,4,2,1.5,0] + degenerate ones
KAT_A = np.array([[0, 0, 0, 4, 2, 1.5, 0]], dtype=np.float32)
KAT_B = np.array([
    [0, 0, 0, 4, 2, 1.5, 0],             # identical
    [1, 0.5, 0, 4, 2, 1.5, 0.7],         # generic
    [1, 0.5, 0, 4, 2, 1.5, 0.7 + 2 * PI],
    [0, 0, 0, 4, 2, 1.5, PI / 2],        # 90 degree cross
    [0.3, 0.1, 0, 1, 0.5, 1.5, 0.4],     # contained
    [2, 0, 0, 4, 2, 1.5, 0],             # half overlap
    [4.005, 0, 0, 4, 2, 1.5, 0],         # 5 mm gap: margin makes it non-zero
    [4.02, 0, 0, 4, 2, 1.5, 0],          # 2 cm gap
    [4, 0, 0, 4, 2, 1.5, 0],             # edges touching
    [3, 2, 0, 2, 2, 1.5, PI / 4],        # corner touch
    [50, 50, 0, 4, 2, 1.5, 0],           # far
    [1, 0.5, 0, 0, 0, 0, 0],             # zero size
    [0, 0, 0, 4, 2, 1.5, 1e-4],          # almost identical (many near-coincident vertices)
    [1e-3, -1e-3, 0, 4.001, 2.001, 1.5, -1e-4],
    [0, 0, 0, 2, 4, 1.5, PI / 2],        # same rectangle, swapped extents
    [0, 0, 0, 4, 2, 1.5, PI],
    [0, 0, 0, 4, 2, 1.5, -3 * PI],
    [0.5, 0.25, 0, 3, 1, 2.5, 100.0],    # large heading
], dtype=np.float32)


def ref_iou_cpu(ref, a, b):
    out = torch.zeros(a.shape[0], b.shape[0])
    ref.boxes_iou_bev_cpu(torch.from_numpy(a).contiguous(), torch.from_numpy(b).contiguous(), out)
    return out.numpy()


def main():
    ref, roi = R.iou3d_nms_cuda(), R.roiaware_pool3d_cuda()
    assert ref is not None and roi is not None, "run oracle/build_ref.py first"
    out = {}
    sets = {
        "kat": (KAT_A, KAT_B),
        "kat_t": (KAT_B, KAT_A),
        "kat_sq": (KAT_B, KAT_B),
        "car35": synth.clustered_pairs(96, 96, 11, (35, 17.5), synth.KITTI_PRIORS[:1]),
        "ped70": synth.clustered_pairs(96, 96, 12, (70, 35), synth.KITTI_PRIORS[1:2]),
        "mix150": synth.clustered_pairs(96, 96, 13, (150, 75)),
        "dense": synth.dense_overlap(64, 64, seed=14),
        "cfg3iou": synth.cfg3_iou(),
        "cfg1sub": (synth.cfg1()[0][::411], synth.cfg1()[1]),
    }
    for k, (a, b) in sets.items():
        out[f"iou_{k}_a"], out[f"iou_{k}_b"] = a, b
        out[f"iou_{k}_ref"] = ref_iou_cpu(ref, a, b)
    pts, rois = synth.cfg3(n_frames=2, n_points=2048, n_rois=40, seed=21)
    # a few hand-made points against box [0,0,0,4,2,1.5,0.3] in frame 0
    rois[0, 0] = [0, 0, 0, 4, 2, 1.5, 0.3]
    pts[0, :4] = [[0, 0, 0], [1, 1, 0.74], [1, 1, 0.76], [5, 5, 0]]
    out["pib_pts"], out["pib_boxes"] = pts, rois
    masks = []
    for f in range(pts.shape[0]):
        m = torch.zeros(rois.shape[1], pts.shape[1], dtype=torch.int32)
        roi.points_in_boxes_cpu(torch.from_numpy(rois[f]).contiguous(), torch.from_numpy(pts[f]).contiguous(), m)
        masks.append(m.numpy())
    out["pib_ref_mask"] = np.stack(masks)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("ref") or k.endswith("mask")})


if __name__ == "__main__":
    main()
