#!/usr/bin/env python
"""Generate golden_gpu.npz from the UNMODIFIED reference CUDA kernels (oracle/_ref, compiled for sm_100a
by oracle/build_ref.py) running on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py'      # writes gpurun_out/golden_gpu.npz
    cp gpurun_out/golden_gpu.npz tests/golden/

Contents (inputs are regenerated from the same seeds as make_golden_cpu.py, so only outputs are stored
for the big sets): boxes_iou_bev_gpu, boxes_overlap_bev_gpu, the reference Python boxes_iou3d_gpu around
its own kernel, nms_gpu / nms_normal_gpu keep lists with the torch sort order that produced them, and
points_in_boxes_gpu.  These pin the CUDA flavor of the oracle in the CPU-only test suite.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lidardetection_b200 import synth  # noqa: E402
from oracle import ref_loader as R  # noqa: E402
import make_golden_cpu as G  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ref, roi = R.iou3d_nms_cuda(), R.roiaware_pool3d_cuda()
    assert ref is not None and roi is not None, "oracle/_ref missing (build it in the dev container first)"
    out = {"gpu_name": np.array(torch.cuda.get_device_name(0))}

    def iou_bev(a, b):
        o = torch.zeros((a.shape[0], b.shape[0]), device=dev)
        ref.boxes_iou_bev_gpu(a, b, o)
        return o

    def overlap(a, b):
        o = torch.zeros((a.shape[0], b.shape[0]), device=dev)
        ref.boxes_overlap_bev_gpu(a, b, o)
        return o

    def iou3d(a, b):  # iou3d_nms_utils.py:48-81 verbatim semantics
        ahmax = (a[:, 2] + a[:, 5] / 2).view(-1, 1)
        ahmin = (a[:, 2] - a[:, 5] / 2).view(-1, 1)
        bhmax = (b[:, 2] + b[:, 5] / 2).view(1, -1)
        bhmin = (b[:, 2] - b[:, 5] / 2).view(1, -1)
        ov = overlap(a, b)
        oh = torch.clamp(torch.min(ahmax, bhmax) - torch.max(ahmin, bhmin), min=0)
        o3d = ov * oh
        va = (a[:, 3] * a[:, 4] * a[:, 5]).view(-1, 1)
        vb = (b[:, 3] * b[:, 4] * b[:, 5]).view(1, -1)
        return o3d / torch.clamp(va + vb - o3d, min=1e-6)

    sets = {
        "kat": (G.KAT_A, G.KAT_B),
        "kat_t": (G.KAT_B, G.KAT_A),
        "kat_sq": (G.KAT_B, G.KAT_B),
        "car35": synth.clustered_pairs(96, 96, 11, (35, 17.5), synth.KITTI_PRIORS[:1]),
        "ped70": synth.clustered_pairs(96, 96, 12, (70, 35), synth.KITTI_PRIORS[1:2]),
        "mix150": synth.clustered_pairs(96, 96, 13, (150, 75)),
        "dense": synth.dense_overlap(64, 64, seed=14),
        "cfg3iou": synth.cfg3_iou(),
        "cfg1sub": (synth.cfg1()[0][::411], synth.cfg1()[1]),
    }
    for k, (a, b) in sets.items():
        ta, tb = torch.from_numpy(a).to(dev).contiguous(), torch.from_numpy(b).to(dev).contiguous()
        out[f"iou_{k}_a"], out[f"iou_{k}_b"] = a, b
        out[f"iou_{k}_bev"] = iou_bev(ta, tb).cpu().numpy()
        out[f"iou_{k}_overlap"] = overlap(ta, tb).cpu().numpy()
        out[f"iou_{k}_iou3d"] = iou3d(ta, tb).cpu().numpy()

    # NMS: 3 clustered frames of 1024 boxes, two thresholds, rotated + normal
    boxes, scores = synth.nms_frames(3, 1024, seed=31)
    out["nms_boxes"], out["nms_scores"] = boxes, scores
    for f in range(3):
        tb, ts = torch.from_numpy(boxes[f]).to(dev), torch.from_numpy(scores[f]).to(dev)
        order = ts.sort(0, descending=True)[1]
        out[f"nms_order_{f}"] = order.cpu().numpy()
        b = tb[order].contiguous()
        for thr in (0.01, 0.1, 0.7):
            for normal in (0, 1):
                keep = torch.LongTensor(b.size(0))
                n = (ref.nms_normal_gpu if normal else ref.nms_gpu)(b, keep, thr)
                out[f"nms_keep_{f}_{thr}_{normal}"] = order[keep[:n].to(dev)].cpu().numpy()

    # points in boxes (GPU form, margin 1e-5, first hit)
    pts, rois = synth.cfg3(n_frames=2, n_points=4096, n_rois=60, seed=41)
    rois[0, 0] = [0, 0, 0, 4, 2, 1.5, 0.3]
    pts[0, :4] = [[0, 0, 0], [1, 1, 0.74], [1, 1, 0.76], [5, 5, 0]]
    rois[1, 5] = rois[1, 4]  # duplicate box: lowest index must win
    out["pib_pts"], out["pib_boxes"] = pts, rois
    o = torch.full((2, pts.shape[1]), -1, dtype=torch.int32, device=dev)
    roi.points_in_boxes_gpu(torch.from_numpy(rois).to(dev).contiguous(), torch.from_numpy(pts).to(dev).contiguous(), o)
    out["pib_idx"] = o.cpu().numpy()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "golden_gpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
