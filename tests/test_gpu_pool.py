"""GPU parity tests of the "next" row SURVEY 8f-3: RoI-aware voxel pooling (forward / backward) and RoI point pooling
through the Python drop-in modules -> ctypes -> C ABI, against
  (1) the oracle restatement (oracle/lg_oracle.c) on the same seeded inputs,
  (2) the committed golden vectors of the reference CUDA kernels (tests/golden/golden_gpu_pool.npz),
  (3) when oracle/_ref travelled with the snapshot, the reference CUDA kernels run live on this GPU.
Bar: voxel lists, counts, argmax, sample lists and every forward value bit-exact (the forward is a selection plus, for
avg pooling, a sum in list order and one IEEE division).  The backward accumulates with float atomics in both
implementations (unordered sums): 1e-5 relative to the largest gradient.
"""
import os
import sys

import numpy as np
import pytest
import torch

from lidardetection_b200 import _lib, synth
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU
from lidardetection_b200.ops.roipoint_pool3d import roipoint_pool3d_utils as RU
from oracle import lg_oracle as O
from oracle import ref_loader as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_gpu_pool as GP  # noqa: E402

pytestmark = pytest.mark.gpu


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda:0")


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def check_forward(rois, pts, feat, out, max_pts, method):
    pooled, argmax, pidx = PU.roiaware_pool3d_forward(cu(rois), cu(pts), cu(feat), out, max_pts, method)
    wp, wa, wi = O.roiaware_pool3d_forward(rois, pts, feat, out, max_pts, method)
    assert np.array_equal(pidx.cpu().numpy(), wi), "voxel lists differ"
    if method == "max":
        assert np.array_equal(argmax.cpu().numpy(), wa), "argmax differs"
    else:
        assert argmax.shape == pooled.shape and int(argmax.abs().sum()) == 0  # zero-filled, as the reference allocates it (roiaware_pool3d_utils.py:85)
    assert np.array_equal(bits(pooled.cpu().numpy()), bits(wp)), "pooled features differ"
    return pooled, argmax, pidx


# (n_points, n_rois, channels, out_size, max_pts)
SHAPES = [
    (4096, 16, 5, 4, 8),            # tiny voxel grid: lists overflow max_pts
    (16384, 128, 128, 12, 128),     # Part-A2 (partA2_head.py:53-56, 138-143): the RPN-feature call
    (16384, 128, 4, 12, 128),       # ... and the part-location call
    (40000, 9, 33, (3, 5, 7), 16),  # three passes of 16384 points per CTA, odd channel count
    (5000, 3, 1, 24, 4),            # 13 824 voxels: counters in global memory
    (100, 2, 8, 1, 128),            # a single voxel per box
    (3000, 4, 6, 5, 1),             # max_pts 1: lists can hold nothing (kernel.cu:85)
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("method", ["max", "avg"])
def test_roiaware_forward_equals_oracle(shape, method):
    m, n, c, out, mp = shape
    pts, rois, feat = synth.pool_case(m, n, c, seed=1000 + m + n)
    rois = rois.copy()
    rois[0, 3:5] *= 3.0
    check_forward(rois, pts, feat, out, mp, method)


def test_roiaware_empty_inputs():
    pts, rois, feat = synth.pool_case(512, 4, 3, seed=5)
    p, a, i = PU.roiaware_pool3d_forward(cu(rois[:0]), cu(pts), cu(feat), 4, 8, "max")
    assert p.shape == (0, 4, 4, 4, 3) and a.shape == p.shape and i.shape == (0, 4, 4, 4, 8)
    p, a, i = PU.roiaware_pool3d_forward(cu(rois), cu(pts[:0]), cu(feat[:0]), 4, 8, "max")
    assert int(i.abs().sum()) == 0 and float(p.abs().sum()) == 0.0 and bool((a == -1).all())
    with pytest.raises(_lib.LidarGeomError):
        PU.roiaware_pool3d_forward(cu(rois), cu(pts), cu(feat), 256, 8, "max")


def test_roiaware_special_values_and_degenerate_boxes():
    """NaN / -inf features never win a max (kernel.cu:131-139: `>` against -inf); zero-length box: x_res = 0"""
    for name, pts, rois, feat, out, mp, s in GP.cases():
        feat = feat.copy()
        feat[::7, 0] = np.nan
        feat[::5, 1] = -np.inf
        feat[::3, 2] = np.inf
        pts_t, rois_t, feat_t = cu(pts), cu(rois), cu(feat)
        pooled, argmax, pidx = PU.roiaware_pool3d_forward(rois_t, pts_t, feat_t, out, mp, "max")
        wp, wa, wi = O.roiaware_pool3d_forward(rois, pts, feat, out, mp, "max")
        assert np.array_equal(pidx.cpu().numpy(), wi) and np.array_equal(argmax.cpu().numpy(), wa)
        assert np.array_equal(bits(pooled.cpu().numpy()), bits(wp))
        assert int(wi[-1, ..., 0].sum()) == 5  # the points on the zero-length box's axis


@pytest.mark.parametrize("method", ["max", "avg"])
def test_roiaware_backward_and_autograd(method):
    pts, rois, feat = synth.pool_case(8192, 24, 16, seed=77)
    layer = PU.RoIAwarePool3d(out_size=6, max_pts_each_voxel=32)
    f = cu(feat).requires_grad_(True)
    pooled = layer(cu(rois), cu(pts), f, pool_method=method)
    g = np.random.default_rng(3).standard_normal(tuple(pooled.shape)).astype(np.float32)
    pooled.backward(cu(g))
    wp, wa, wi = O.roiaware_pool3d_forward(rois, pts, feat, 6, 32, method)
    want = O.roiaware_pool3d_backward(wi, wa, g, pts.shape[0], method)
    got = f.grad.cpu().numpy()
    assert got.shape == want.shape
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-5 * scale
    assert np.array_equal(got == 0, want == 0)  # points outside every box get exactly no gradient


def test_roipoint_forward_equals_oracle():
    for (n, m, c, s, seed) in [(16384, 128, 128, 512, 1), (5000, 7, 3, 33, 2), (700, 5, 0, 16, 3), (20000, 4, 5, 2048, 4)]:
        B = 2
        frames = [synth.pool_case(n, m, max(c, 1), seed=seed * 10 + b) for b in range(B)]
        pts = np.stack([f[0] for f in frames])
        boxes = np.stack([f[1] for f in frames])
        feat = np.stack([f[2][:, :c] for f in frames])
        boxes[0, 0, 3:5] *= 6.0  # > S points inside
        boxes[1, 1, 0:2] += 500.0  # an empty box
        layer = RU.RoIPointPool3d(num_sampled_points=s, pool_extra_width=(0.2, 0.2, 0.2))
        pooled, flag = layer(cu(pts), cu(feat), cu(boxes))
        big = boxes.copy()
        big[..., 3:6] += np.float32(0.2)
        wp, wf = O.roipoint_pool3d_forward(pts, feat, big, s)
        assert np.array_equal(flag.cpu().numpy(), wf)
        assert wf[1, 1] == 1 and wf[0, 0] == 0
        assert np.array_equal(bits(pooled.cpu().numpy()), bits(wp))


def test_pool_golden_vectors_of_reference_kernels():
    p = os.path.join(HERE, "golden", "golden_gpu_pool.npz")
    if not os.path.exists(p):
        pytest.skip("golden_gpu_pool.npz not generated yet (tests/golden/make_golden_gpu_pool.py via gpurun)")
    g = np.load(p)
    for name, pts, rois, feat, out, mp, s in GP.cases():
        for method in ("max", "avg"):
            pooled, argmax, pidx = PU.roiaware_pool3d_forward(cu(rois), cu(pts), cu(feat), out, mp, method)
            assert np.array_equal(bits(pooled.cpu().numpy()), bits(g[f"{name}_{method}_pooled"])), (name, method)
            assert np.array_equal(pidx.cpu().numpy(), g[f"{name}_pts_idx"])
            if method == "max":
                assert np.array_equal(argmax.cpu().numpy(), g[f"{name}_argmax"])
            go = np.random.default_rng(7).standard_normal(tuple(pooled.shape)).astype(np.float32)
            gi = PU.roiaware_pool3d_backward(pidx, argmax, cu(go), pts.shape[0], method).cpu().numpy()
            want = g[f"{name}_{method}_grad_in"]
            assert np.abs(gi - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max())), (name, method)
        big = rois.copy()
        big[:, 3:6] += np.float32(0.2)
        pf, fl = RU.roipoint_pool3d_forward(cu(pts[None]), cu(big[None]), cu(feat[None]), s)
        assert np.array_equal(fl.cpu().numpy(), g[f"{name}_rp_flag"])
        assert np.array_equal(bits(pf.cpu().numpy()), bits(g[f"{name}_rp_pooled"]))


def test_pool_equals_live_reference_kernels():
    roi, rpp = R.roiaware_pool3d_cuda(), R.roipoint_pool3d_cuda()
    if roi is None or rpp is None:
        pytest.skip("oracle/_ref not built in this snapshot")
    pts, rois, feat = synth.pool_case(16384, 64, 32, seed=4242)
    tr, tp, tf = cu(rois), cu(pts), cu(feat)
    n, c, out, mp = rois.shape[0], feat.shape[1], (12, 12, 12), 128
    for method, mi in (("max", 0), ("avg", 1)):
        pooled = tf.new_zeros((n, *out, c))
        argmax = tf.new_zeros((n, *out, c), dtype=torch.int)
        pidx = tf.new_zeros((n, *out, mp), dtype=torch.int)
        roi.forward(tr, tp, tf, argmax, pidx, pooled, mi)
        gp, ga, gi = PU.roiaware_pool3d_forward(tr, tp, tf, out, mp, method)
        assert torch.equal(gi, pidx)
        assert torch.equal(gp.view(torch.int32), pooled.view(torch.int32))
        if mi == 0:
            assert torch.equal(ga, argmax)
        g = torch.randn_like(pooled)
        want = g.new_zeros((pts.shape[0], c))
        roi.backward(pidx, argmax, g, want, mi)
        got = PU.roiaware_pool3d_backward(gi, ga, g, pts.shape[0], method)
        assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    tb = tr.clone()[None]
    tb[..., 3:6] += 0.2
    pf = tf.new_zeros((1, n, 512, 3 + c))
    fl = tf.new_zeros((1, n)).int()
    rpp.forward(tp[None].contiguous(), tb.contiguous(), tf[None].contiguous(), pf, fl)
    gpf, gfl = RU.roipoint_pool3d_forward(tp[None], tb, tf[None], 512)
    assert torch.equal(gfl, fl) and torch.equal(gpf.view(torch.int32), pf.view(torch.int32))
