#!/usr/bin/env python
"""Developer diagnostic (run on a B200 through gpurun): differential statistics of liblidargeom against
the oracle restatement and against the compiled reference CUDA kernels (oracle/_ref, "tier A"), plus
quick CUDA-event timings of both.  Writes gpurun_out/gpu_check.json.  Not part of the product."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import synth  # noqa: E402
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U  # noqa: E402
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU  # noqa: E402
from lidardetection_b200 import _lib  # noqa: E402
from oracle import lg_oracle as O, ref_loader as R  # noqa: E402

dev = torch.device("cuda:0")
res = {"gpu": torch.cuda.get_device_name(0)}
ref = R.iou3d_nms_cuda()
roi = R.roiaware_pool3d_cuda()
print("reference extensions:", ref is not None, roi is not None)


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def cmp(name, x, y):
    x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
    d = np.abs(x.astype(np.float64) - y.astype(np.float64))
    out = {"n": int(x.size), "bit_mismatch": int((bits(x) != bits(y)).sum()), "max_abs": float(d.max()) if d.size else 0.0,
           "gt1e-5": int((d > 1e-5).sum()), "gt1e-6": int((d > 1e-6).sum()), "nonzero": int((y != 0).sum())}
    print(f"  {name:46s} {out}")
    return out


def ev_time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def ref_iou_bev(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), device=dev)
    ref.boxes_iou_bev_gpu(a.contiguous(), b.contiguous(), out)
    return out


def ref_overlap(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), device=dev)
    ref.boxes_overlap_bev_gpu(a.contiguous(), b.contiguous(), out)
    return out


def ref_iou3d(a, b):
    """the reference's Python boxes_iou3d_gpu (iou3d_nms_utils.py:48-81) around its own overlap kernel"""
    ahmax = (a[:, 2] + a[:, 5] / 2).view(-1, 1)
    ahmin = (a[:, 2] - a[:, 5] / 2).view(-1, 1)
    bhmax = (b[:, 2] + b[:, 5] / 2).view(1, -1)
    bhmin = (b[:, 2] - b[:, 5] / 2).view(1, -1)
    ov = ref_overlap(a, b)
    max_of_min = torch.max(ahmin, bhmin)
    min_of_max = torch.min(ahmax, bhmax)
    oh = torch.clamp(min_of_max - max_of_min, min=0)
    o3d = ov * oh
    va = (a[:, 3] * a[:, 4] * a[:, 5]).view(-1, 1)
    vb = (b[:, 3] * b[:, 4] * b[:, 5]).view(1, -1)
    return o3d / torch.clamp(va + vb - o3d, min=1e-6)


def ref_nms(boxes, scores, thresh, normal=False):
    order = scores.sort(0, descending=True)[1]
    b = boxes[order].contiguous()
    keep = torch.LongTensor(b.size(0))
    n = (ref.nms_normal_gpu if normal else ref.nms_gpu)(b, keep, thresh)
    return order[keep[:n].to(dev)].contiguous()


# ------------------------------------------------------------------ IoU differential
sets = {
    "car35": synth.clustered_pairs(400, 400, 1, (35, 17.5), synth.KITTI_PRIORS[:1]),
    "ped70": synth.clustered_pairs(400, 400, 2, (70, 35), synth.KITTI_PRIORS[1:2]),
    "mix150": synth.clustered_pairs(400, 400, 3, (150, 75)),
    "dense": synth.dense_overlap(400, 400),
    "cfg3iou": synth.cfg3_iou(),
    "cfg1sub": (synth.cfg1()[0][::37], synth.cfg1()[1]),
}
res["iou"] = {}
for name, (a, b) in sets.items():
    print(name, a.shape, b.shape)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    ours = U.boxes_iou_bev(ta, tb).cpu().numpy()
    ours_ov = U.boxes_overlap_bev(ta, tb).cpu().numpy()
    ours_3d = U.boxes_iou3d_gpu(ta, tb).cpu().numpy()
    ours_strict = U._iou_call("lg_boxes_iou_bev", ta, tb, flags=_lib.LG_FLAG_STRICT_FP32).cpu().numpy()
    o1 = O.boxes_iou_bev(a, b, 1)
    o0 = O.boxes_iou_bev(a, b, 0)
    r = {}
    r["ours_vs_oracle_cuda"] = cmp("ours vs oracle(cuda flavor)", ours, o1)
    r["ours_strict_vs_oracle_cpu"] = cmp("ours(strict) vs oracle(cpu flavor)", ours_strict, o0)
    r["ours_ov_vs_oracle"] = cmp("ours overlap vs oracle(cuda)", ours_ov, O.boxes_overlap_bev(a, b, 1))
    r["ours_3d_vs_oracle"] = cmp("ours iou3d vs oracle(cuda)", ours_3d, O.boxes_iou3d(a, b, 1))
    if ref is not None:
        ra = ref_iou_bev(ta, tb).cpu().numpy()
        r["ours_vs_refgpu"] = cmp("ours vs reference GPU kernel", ours, ra)
        r["oracle_cuda_vs_refgpu"] = cmp("oracle(cuda flavor) vs reference GPU", o1, ra)
        r["oracle_cpu_vs_refgpu"] = cmp("oracle(cpu flavor) vs reference GPU", o0, ra)
        r["ours_ov_vs_refgpu"] = cmp("ours overlap vs reference GPU", ours_ov, ref_overlap(ta, tb).cpu().numpy())
        r["ours_3d_vs_refgpu"] = cmp("ours iou3d vs reference python+GPU", ours_3d, ref_iou3d(ta, tb).cpu().numpy())
    res["iou"][name] = r

# ------------------------------------------------------------------ NMS
res["nms"] = {}
boxes, scores = synth.cfg2(n_frames=4, n_boxes=4096)
for normal in (False, True):
    for f in range(4):
        tb_, ts_ = torch.from_numpy(boxes[f]).to(dev), torch.from_numpy(scores[f]).to(dev)
        thresh = 0.01 if f < 2 else 0.5
        ours = (U.nms_normal_gpu if normal else U.nms_gpu)(tb_, ts_, thresh)[0].cpu().numpy()
        order = ts_.sort(0, descending=True)[1].cpu().numpy()
        orc = O.nms(boxes[f], scores[f], thresh, normal=normal, flavor=1, order=order)
        entry = {"kept_ours": int(len(ours)), "kept_oracle": int(len(orc)), "equal_oracle": bool(np.array_equal(ours, orc))}
        if ref is not None:
            rr = ref_nms(tb_, ts_, thresh, normal).cpu().numpy()
            entry.update({"kept_ref": int(len(rr)), "equal_ref": bool(np.array_equal(ours, rr)),
                          "oracle_equal_ref": bool(np.array_equal(orc, rr))})
        print("nms", "normal" if normal else "rotated", f, thresh, entry)
        res["nms"][f"{'normal' if normal else 'rot'}_{f}"] = entry
# batched
tb_, ts_ = torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev)
keep, num = U.nms_gpu_batched(tb_, ts_, 0.01)
ok = True
for f in range(4):
    single = U.nms_gpu(tb_[f], ts_[f], 0.01)[0]
    ok &= bool(torch.equal(single, keep[f, : int(num[f])]))
print("batched == single:", ok)
res["nms"]["batched_equals_single"] = ok

# ------------------------------------------------------------------ points
res["points"] = {}
pts, rois = synth.cfg3(n_frames=4)
tp, tr = torch.from_numpy(pts).to(dev), torch.from_numpy(rois).to(dev)
ours = PU.points_in_boxes_gpu(tp, tr).cpu().numpy()
orc = O.points_in_boxes_idx(pts, rois, 1)
entry = {"mismatch_oracle": int((ours != orc).sum()), "inside": int((orc >= 0).sum())}
if roi is not None:
    out = torch.full((4, pts.shape[1]), -1, dtype=torch.int32, device=dev)
    roi.points_in_boxes_gpu(tr.contiguous(), tp.contiguous(), out)
    entry["mismatch_ref"] = int((ours != out.cpu().numpy()).sum())
    entry["oracle_mismatch_ref"] = int((orc != out.cpu().numpy()).sum())
mk = PU.points_in_boxes_cpu(pts[0], rois[0])
entry["mask_mismatch_oracle_cpu"] = int((mk != O.points_in_boxes_mask(pts[0], rois[0], 1e-2, 0)).sum())
print("points", entry)
res["points"] = entry

# ------------------------------------------------------------------ timings (ms, median of 10)
res["time_ms"] = {}


def timed(name, fn, pairs=None):
    t = ev_time(fn)
    res["time_ms"][name] = t
    extra = f"  {pairs / t / 1e6:.2f} Gpairs/s" if pairs else ""
    print(f"time {name:40s} {t:9.3f} ms{extra}")


a, b = synth.dense_overlap(8192, 8192)
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
timed("ours iou_bev dense 8192^2", lambda: U.boxes_iou_bev(ta, tb), 8192 * 8192)
if ref is not None:
    timed("ref  iou_bev dense 8192^2", lambda: ref_iou_bev(ta, tb), 8192 * 8192)
a, b = synth.cfg1()
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
timed("ours iou_bev cfg1 321408x20", lambda: U.boxes_iou_bev(ta, tb), a.shape[0] * 20)
timed("ours iou3d   cfg1 321408x20", lambda: U.boxes_iou3d_gpu(ta, tb), a.shape[0] * 20)
if ref is not None:
    timed("ref  iou_bev cfg1 321408x20", lambda: ref_iou_bev(ta, tb), a.shape[0] * 20)
    timed("ref  iou3d   cfg1 321408x20", lambda: ref_iou3d(ta, tb), a.shape[0] * 20)
a, b = synth.cfg4(n=32768)
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
timed("ours iou3d cfg4-like 32768^2", lambda: U.boxes_iou3d_gpu(ta, tb), 32768 * 32768)
if ref is not None:
    timed("ref  overlap cfg4-like 32768^2", lambda: ref_overlap(ta, tb), 32768 * 32768)
boxes, scores = synth.cfg2(n_frames=16, n_boxes=4096)
tb_, ts_ = torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev)
timed("ours nms batched 16x4096", lambda: U.nms_gpu_batched(tb_, ts_, 0.01))
timed("ours nms loop    16x4096", lambda: [U.nms_gpu(tb_[f], ts_[f], 0.01) for f in range(16)])
if ref is not None:
    timed("ref  nms loop    16x4096", lambda: [ref_nms(tb_[f], ts_[f], 0.01) for f in range(16)])
pts, rois = synth.cfg3(n_frames=64)
tp, tr = torch.from_numpy(pts).to(dev), torch.from_numpy(rois).to(dev)
timed("ours points 64x16384x100", lambda: PU.points_in_boxes_gpu(tp, tr))
if roi is not None:
    out = torch.full((64, pts.shape[1]), -1, dtype=torch.int32, device=dev)
    timed("ref  points 64x16384x100", lambda: roi.points_in_boxes_gpu(tr, tp, out))

# post-processing front end: 64 frames x 70,400 candidates (SECOND KITTI head), score_thresh 0.1, 4096 -> 500
from lidardetection_b200 import model_nms_utils as MU  # noqa: E402

cfgp = {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.01, "NMS_PRE_MAXSIZE": 4096, "NMS_POST_MAXSIZE": 500}
bx, sc = synth.cfg2(n_frames=64, n_boxes=4096)
rr = np.random.default_rng(1)
allb = np.concatenate([bx, synth.gt_boxes(64 * (70400 - 4096), 9).reshape(64, -1, 7)], 1)
alls = np.concatenate([sc, rr.uniform(0.0, 0.099, (64, 70400 - 4096)).astype(np.float32)], 1)
perm = rr.permutation(70400)
tb_, ts_ = torch.from_numpy(allb[:, perm]).to(dev), torch.from_numpy(alls[:, perm]).to(dev)
timed("ours post-proc batched 64x70400", lambda: MU.class_agnostic_nms_batched(ts_, tb_, cfgp, score_thresh=0.1))
timed("ours post-proc per-frame loop (reference structure)", lambda: [MU.class_agnostic_nms(ts_[f], tb_[f], cfgp, score_thresh=0.1) for f in range(64)])
selb, numb, _ = MU.class_agnostic_nms_batched(ts_, tb_, cfgp, score_thresh=0.1)
okp = all(torch.equal(MU.class_agnostic_nms(ts_[f], tb_[f], cfgp, score_thresh=0.1)[0], selb[f, : int(numb[f])]) for f in range(0, 64, 7))
print("post-proc batched == per-frame:", okp)
res["post_processing_batched_equals_per_frame"] = bool(okp)

# fused row / column maxima against matrix + torch.max
a, b = synth.cfg1()
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)


def unfused(x, y):
    m_ = U.boxes_iou3d_gpu(x, y)
    return m_.max(1), m_.max(0)


timed("ours iou3d + torch.max both axes, cfg1 321408x20", lambda: unfused(ta, tb))
timed("ours boxes_iou_max both axes,     cfg1 321408x20", lambda: U.boxes_iou_max(ta, tb, rows=True, cols=True))
a, b = synth.cfg4(200000)
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
timed("ours boxes_iou_max both axes, 200k x 200k (160 GB matrix never built)", lambda: U.boxes_iou_max(ta, tb, rows=True, cols=True), 4e10)

# RoI-aware pooling / RoI point pooling against the reference kernels, one Part-A2 / PointRCNN frame
from lidardetection_b200.ops.roipoint_pool3d import roipoint_pool3d_utils as RU  # noqa: E402

rpp = R.roipoint_pool3d_cuda()
pts1, rois1, feat1 = synth.pool_case(16384, 128, 128)
tp1, tr1, tf1 = (torch.from_numpy(x).to(dev) for x in (pts1, rois1, feat1))
tf4 = tf1[:, :4].contiguous()
timed("ours roiaware fwd max C=128 (128 rois, 12^3)", lambda: PU.roiaware_pool3d_forward(tr1, tp1, tf1, 12, 128, "max"))
timed("ours roiaware fwd avg C=4   (128 rois, 12^3)", lambda: PU.roiaware_pool3d_forward(tr1, tp1, tf4, 12, 128, "avg"))
po, am, pi = PU.roiaware_pool3d_forward(tr1, tp1, tf1, 12, 128, "max")
go = torch.randn_like(po)
timed("ours roiaware bwd max C=128", lambda: PU.roiaware_pool3d_backward(pi, am, go, 16384, "max"))
timed("ours roiaware bwd avg C=128", lambda: PU.roiaware_pool3d_backward(pi, am, go, 16384, "avg"))
tb1 = tr1[None].contiguous()
timed("ours roipoint fwd 1x128x512x(3+128)", lambda: RU.roipoint_pool3d_forward(tp1[None], tb1, tf1[None], 512))
if roi is not None and rpp is not None:
    def ref_roiaware(f, method):
        c = f.shape[1]
        pooled = f.new_zeros((128, 12, 12, 12, c))
        argmax = f.new_zeros((128, 12, 12, 12, c), dtype=torch.int)
        pidx = f.new_zeros((128, 12, 12, 12, 128), dtype=torch.int)
        roi.forward(tr1, tp1, f, argmax, pidx, pooled, method)
        return pooled, argmax, pidx

    def ref_roiaware_bwd(method):
        gi = go.new_zeros((16384, 128))
        roi.backward(pi, am, go, gi, method)
        return gi

    def ref_roipoint():
        pf = tf1.new_zeros((1, 128, 512, 131))
        fl = tf1.new_zeros((1, 128)).int()
        rpp.forward(tp1[None].contiguous(), tb1, tf1[None].contiguous(), pf, fl)
        return pf, fl

    timed("ref  roiaware fwd max C=128 (128 rois, 12^3)", lambda: ref_roiaware(tf1, 0))
    timed("ref  roiaware fwd avg C=4   (128 rois, 12^3)", lambda: ref_roiaware(tf4, 1))
    timed("ref  roiaware bwd max C=128", lambda: ref_roiaware_bwd(0))
    timed("ref  roiaware bwd avg C=128", lambda: ref_roiaware_bwd(1))
    timed("ref  roipoint fwd 1x128x512x(3+128)", ref_roipoint)
    rp, ra, ri = ref_roiaware(tf1, 0)
    okr = bool(torch.equal(ri, pi) and torch.equal(ra, am) and torch.equal(rp.view(torch.int32), po.view(torch.int32)))
    print("roiaware forward == reference kernels (bit-exact):", okr)
    res["roiaware_equals_reference"] = okr

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
    json.dump(res, f, indent=1)
print("done")
