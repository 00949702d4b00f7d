"""Route B of INTEGRATION.md, executed: the reference's OWN, UNMODIFIED Python wrappers
(pcdet/ops/iou3d_nms/iou3d_nms_utils.py, pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py -- byte-compiled by
oracle/build_ref.py into oracle/_ref/py/, because /root/reference does not exist on the GPU box) run on top of the ctypes
shims of lidardetection_b200/compat/ (the C ABI under the extension modules' names, iou3d_nms_api.cpp:11-17,
roiaware_pool3d.cpp:172-177), and are compared with the same Python on top of the reference's compiled extensions
(oracle/_ref/*.so): every function of the two modules, same inputs, on this GPU.
"""
import numpy as np
import pytest
import torch

from lidardetection_b200 import compat, synth
from lidardetection_b200.compat import iou3d_nms_cuda as shim_iou, roiaware_pool3d_cuda as shim_roi
from oracle import ref_loader as R


def test_shims_export_the_extension_interfaces():
    """CPU tier: the names the reference's pybind modules define (iou3d_nms_api.cpp:11-17, roiaware_pool3d.cpp:172-177)"""
    for name in ("boxes_overlap_bev_gpu", "boxes_iou_bev_gpu", "nms_gpu", "nms_normal_gpu", "boxes_iou_bev_cpu"):
        assert callable(getattr(shim_iou, name))
    for name in ("forward", "backward", "points_in_boxes_gpu", "points_in_boxes_cpu"):
        assert callable(getattr(shim_roi, name))
    import sys

    a, b = compat.install("pcdet_rb_probe")
    assert sys.modules["pcdet_rb_probe.ops.iou3d_nms.iou3d_nms_cuda"] is a is shim_iou
    assert sys.modules["pcdet_rb_probe.ops.roiaware_pool3d.roiaware_pool3d_cuda"] is b is shim_roi


def test_reference_python_imports_on_top_of_the_shims():
    """CPU tier: the unmodified reference modules import against the shims and expose their public functions"""
    m = R.reference_python("iou3d_nms_utils", shim_iou, "iou3d_nms_cuda", "pcdet_rb_cpu")
    if m is None:
        pytest.skip("oracle/_ref/py (byte-compiled reference wrappers) not built")
    for name in ("boxes_bev_iou_cpu", "boxes_iou_bev", "boxes_iou3d_gpu", "nms_gpu", "nms_normal_gpu"):
        assert callable(getattr(m, name))
    assert m.iou3d_nms_cuda is shim_iou
    p = R.reference_python("roiaware_pool3d_utils", shim_roi, "roiaware_pool3d_cuda", "pcdet_rb_cpu")
    assert p.roiaware_pool3d_cuda is shim_roi and callable(p.points_in_boxes_gpu) and callable(p.points_in_boxes_cpu)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


@pytest.mark.gpu
def test_unmodified_reference_python_over_shims_equals_reference_python_over_its_own_extensions():
    if not R.available():
        pytest.skip("oracle/_ref (compiled reference) did not travel with this snapshot")
    ours_u = R.reference_python("iou3d_nms_utils", shim_iou, "iou3d_nms_cuda", "pcdet_rb_ours")
    ref_u = R.reference_python("iou3d_nms_utils", R.iou3d_nms_cuda(), "iou3d_nms_cuda", "pcdet_rb_ref")
    ours_p = R.reference_python("roiaware_pool3d_utils", shim_roi, "roiaware_pool3d_cuda", "pcdet_rb_ours")
    ref_p = R.reference_python("roiaware_pool3d_utils", R.roiaware_pool3d_cuda(), "roiaware_pool3d_cuda", "pcdet_rb_ref")
    if ours_u is None or ours_p is None:
        pytest.skip("oracle/_ref/py (byte-compiled reference wrappers) not built")
    assert ours_u is not ref_u and ours_u.iou3d_nms_cuda is shim_iou and ref_u.iou3d_nms_cuda is R.iou3d_nms_cuda()

    # ---- iou3d_nms_utils.py:31-81: boxes_iou_bev, boxes_iou3d_gpu
    a, b = synth.clustered_pairs(700, 450, 31, (40, 10))
    ta, tb = cu(a), cu(b)
    for fn in ("boxes_iou_bev", "boxes_iou3d_gpu"):
        got, want = getattr(ours_u, fn)(ta, tb), getattr(ref_u, fn)(ta, tb)
        assert got.shape == want.shape == (700, 450)
        assert float((got - want).abs().max()) <= 1e-5
        assert int((bits(got) != bits(want)).sum()) <= 700 * 450 // 2000, fn  # vertex-order ties only
        assert torch.equal(got == 0, want == 0)
    # ---- :84-116: nms_gpu (with and without pre_maxsize, config splatted in as **kwargs), nms_normal_gpu
    boxes, scores = synth.cfg2(n_frames=2, n_boxes=4096)
    for f in range(2):
        tbx, ts = cu(boxes[f]), cu(scores[f])
        for thr in (0.01, 0.5):
            g, none = ours_u.nms_gpu(tbx, ts, thr, NMS_TYPE="nms_gpu")
            w, _ = ref_u.nms_gpu(tbx, ts, thr, NMS_TYPE="nms_gpu")
            assert none is None and torch.equal(g, w), (f, thr)
        g, _ = ours_u.nms_gpu(tbx, ts, 0.1, pre_maxsize=1000)
        w, _ = ref_u.nms_gpu(tbx, ts, 0.1, pre_maxsize=1000)
        assert torch.equal(g, w)
        g, _ = ours_u.nms_normal_gpu(tbx, ts, 0.3)
        w, _ = ref_u.nms_normal_gpu(tbx, ts, 0.3)
        assert torch.equal(g, w)
    # ---- :12-28: boxes_bev_iou_cpu, numpy in / numpy out, the reference's compiled CPU function as the other side
    a2, b2 = synth.clustered_pairs(300, 200, 32, (70, 35), synth.KITTI_PRIORS[1:2])  # pedestrians at range: the stress case
    got, want = ours_u.boxes_bev_iou_cpu(a2, b2), ref_u.boxes_bev_iou_cpu(a2, b2)
    assert isinstance(got, np.ndarray) and got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5 and np.array_equal(got == 0, want == 0)
    assert int((got.view(np.uint32) != want.view(np.uint32)).sum()) <= 30
    # ---- roiaware_pool3d_utils.py:9-41: points_in_boxes_gpu / points_in_boxes_cpu
    pts, rois = synth.cfg3(n_frames=3, seed=17)
    assert torch.equal(ours_p.points_in_boxes_gpu(cu(pts), cu(rois)), ref_p.points_in_boxes_gpu(cu(pts), cu(rois)))
    gm, wm = ours_p.points_in_boxes_cpu(pts[0], rois[0]), ref_p.points_in_boxes_cpu(pts[0], rois[0])
    assert isinstance(gm, np.ndarray) and gm.dtype == wm.dtype and np.array_equal(gm, wm)
    # ---- :44-107: RoIAwarePool3d forward + backward through the reference's autograd Function
    ppts, prois, pfeat = synth.pool_case(4096, 16, 8, seed=23)
    for method in ("max", "avg"):
        outs = []
        for mod in (ours_p, ref_p):
            feat = cu(pfeat).requires_grad_(True)
            layer = mod.RoIAwarePool3d(out_size=6, max_pts_each_voxel=32)
            pooled = layer(cu(prois), cu(ppts), feat, pool_method=method)
            pooled.sum().backward()
            outs.append((pooled.detach(), feat.grad.detach()))
        assert torch.equal(bits(outs[0][0]), bits(outs[1][0])), method
        assert float((outs[0][1] - outs[1][1]).abs().max()) <= 1e-5
