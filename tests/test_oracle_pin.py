"""CPU-only: pin the oracle (oracle/lg_oracle.c) to the reference.

  * CPU flavor  == the reference's compiled boxes_iou_bev_cpu / points_in_boxes_cpu, bit for bit
                   (golden_cpu.npz, produced by tests/golden/make_golden_cpu.py from oracle/_ref; and live
                   against oracle/_ref when it is present in this checkout).
  * CUDA flavor == the reference's CUDA kernels run on a B200, bit for bit
                   (golden_gpu.npz, produced by tests/golden/make_golden_gpu.py through gpurun).
"""
import os

import numpy as np
import pytest

from lidardetection_b200 import synth
from oracle import lg_oracle as O
from oracle import ref_loader as R

HERE = os.path.dirname(os.path.abspath(__file__))
IOU_SETS = ["kat", "kat_t", "kat_sq", "car35", "ped70", "mix150", "dense", "cfg3iou", "cfg1sub"]


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gcpu():
    return np.load(os.path.join(HERE, "golden", "golden_cpu.npz"))


@pytest.fixture(scope="module")
def ggpu():
    p = os.path.join(HERE, "golden", "golden_gpu.npz")
    if not os.path.exists(p):
        pytest.skip("golden_gpu.npz not generated yet (tests/golden/make_golden_gpu.py via gpurun)")
    return np.load(p)


def test_known_answers_from_survey():
    """SURVEY.md section 8c starter answers (reference CPU oracle), a = [0,0,0,4,2,1.5,0]."""
    a = [0, 0, 0, 4, 2, 1.5, 0]
    pi = np.pi
    cases = [
        ([0, 0, 0, 4, 2, 1.5, 0], 1.0),
        ([1, 0.5, 0, 4, 2, 1.5, 0.7], 0.414507687),
        ([1, 0.5, 0, 4, 2, 1.5, 0.7 + 2 * pi], 0.414507687),
        ([0, 0, 0, 4, 2, 1.5, pi / 2], 0.333333343),
        ([0.3, 0.1, 0, 1, 0.5, 1.5, 0.4], 0.0624999925),
        ([2, 0, 0, 4, 2, 1.5, 0], 0.333333343),
        ([4.005, 0, 0, 4, 2, 1.5, 0], 6.2540517e-4),  # 5 mm gap: the 1 cm margin makes it non-zero
        ([4.02, 0, 0, 4, 2, 1.5, 0], 0.0),
        ([4, 0, 0, 4, 2, 1.5, 0], 0.0),
        ([3, 2, 0, 2, 2, 1.5, pi / 4], 0.0),
        ([50, 50, 0, 4, 2, 1.5, 0], 0.0),
        ([1, 0.5, 0, 0, 0, 0, 0], 0.0),
    ]
    for b, want in cases:
        got = O.iou_bev_pair(a, b, O.FLAVOR_CPU)
        assert got == pytest.approx(want, abs=5e-9, rel=2e-7), (b, got, want)
        if want == 0.0:
            assert got == 0.0 and O.iou_bev_pair(a, b, O.FLAVOR_CUDA) == 0.0  # exact zeros (database sampler tests == 0)
    # IoU is not bit-symmetric
    assert O.iou_bev_pair([1, 0.5, 0, 4, 2, 1.5, 0.7], a, O.FLAVOR_CPU) == pytest.approx(0.414507747, abs=5e-9)


@pytest.mark.parametrize("name", IOU_SETS)
def test_cpu_flavor_bit_exact_vs_reference_cpu(gcpu, name):
    a, b, ref = gcpu[f"iou_{name}_a"], gcpu[f"iou_{name}_b"], gcpu[f"iou_{name}_ref"]
    got = O.boxes_iou_bev(a, b, O.FLAVOR_CPU)
    assert got.shape == ref.shape
    assert np.array_equal(bits(got), bits(ref)), f"{int((bits(got) != bits(ref)).sum())} of {ref.size} differ"


def test_points_cpu_form_bit_exact_vs_reference_cpu(gcpu):
    pts, boxes, ref = gcpu["pib_pts"], gcpu["pib_boxes"], gcpu["pib_ref_mask"]
    for f in range(pts.shape[0]):
        got = O.points_in_boxes_mask(pts[f], boxes[f], 1e-2, O.FLAVOR_CPU)
        assert np.array_equal(got, ref[f])
    # box [0,0,0,4,2,1.5,0.3]: (0,0,0)->1, (1,1,.74)->1, (1,1,.76)->0, (5,5,0)->0
    assert ref[0, 0, :4].tolist() == [1, 1, 0, 0]


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built in this checkout")
def test_cpu_flavor_live_vs_compiled_reference():
    import torch

    ref, roi = R.iou3d_nms_cuda(), R.roiaware_pool3d_cuda()
    for seed, centre, pri in [(101, (20, -10), synth.KITTI_PRIORS), (102, (69, 39), synth.KITTI_PRIORS[1:2]),
                              (103, (-75, 75), synth.WAYMO_PRIORS), (104, (0.3, 0.2), synth.KITTI_PRIORS[:1])]:
        a, b = synth.clustered_pairs(150, 130, seed, centre, pri)
        out = torch.zeros(a.shape[0], b.shape[0])
        ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), out)
        got = O.boxes_iou_bev(a, b, O.FLAVOR_CPU)
        assert np.array_equal(bits(got), bits(out.numpy()))
    pts, rois = synth.cfg3(n_frames=1, n_points=3000, n_rois=50, seed=77)
    m = torch.zeros(50, 3000, dtype=torch.int32)
    roi.points_in_boxes_cpu(torch.from_numpy(rois[0]), torch.from_numpy(pts[0]), m)
    assert np.array_equal(O.points_in_boxes_mask(pts[0], rois[0], 1e-2, O.FLAVOR_CPU), m.numpy())


@pytest.mark.parametrize("name", IOU_SETS)
def test_cuda_flavor_bit_exact_vs_reference_cuda_kernels(ggpu, name):
    a, b = ggpu[f"iou_{name}_a"], ggpu[f"iou_{name}_b"]
    for key, fn in (("bev", O.boxes_iou_bev), ("overlap", O.boxes_overlap_bev), ("iou3d", O.boxes_iou3d)):
        ref = ggpu[f"iou_{name}_{key}"]
        got = fn(a, b, O.FLAVOR_CUDA)
        nbad = int((bits(got) != bits(ref)).sum())
        # Vertex-order ties (glibc atan2f here vs libdevice atan2f on the GPU) may move the last bits of a pair
        # whose polygon has coincident vertices.  The hand-made kat_sq set is built from coincident boxes
        # (identical, swapped-extent, 1e-4 rad apart: 10-12 vertices), so ties are the rule there; the seeded
        # random sets must be bit-identical up to a stray tie.
        allowed = max(2, ref.size // 20) if name.startswith("kat") else max(1, ref.size // 2000)
        assert nbad <= allowed, f"{key}: {nbad} of {ref.size} differ"
        assert np.abs(got - ref).max() <= 1e-6


def test_cuda_flavor_nms_and_points_vs_reference_cuda(ggpu):
    boxes, scores = ggpu["nms_boxes"], ggpu["nms_scores"]
    for f in range(boxes.shape[0]):
        order = ggpu[f"nms_order_{f}"]
        for thr in (0.01, 0.1, 0.7):
            for normal in (0, 1):
                ref = ggpu[f"nms_keep_{f}_{thr}_{normal}"]
                got = O.nms(boxes[f], scores[f], thr, normal=bool(normal), flavor=O.FLAVOR_CUDA, order=order)
                assert np.array_equal(got, ref), (f, thr, normal)
    got = O.points_in_boxes_idx(ggpu["pib_pts"], ggpu["pib_boxes"], O.FLAVOR_CUDA)
    assert np.array_equal(got, ggpu["pib_idx"])
    assert ggpu["pib_idx"][0, :4].tolist() == [0, 0, -1, -1]


def test_nms_full_mask_sweep_equals_lazy_form():
    boxes, scores = synth.nms_frames(1, 300, seed=5)
    order = np.argsort(-scores[0], kind="stable")
    b = boxes[0][order]
    for normal in (False, True):
        for thr in (0.01, 0.3):
            full = O.nms_sorted(b, thr, normal, O.FLAVOR_CUDA, lazy=False)
            lazy = O.nms_sorted(b, thr, normal, O.FLAVOR_CUDA, lazy=True)
            assert np.array_equal(full, lazy)
            assert np.all(np.diff(full) > 0) and full[0] == 0


def test_nms_edge_cases():
    assert O.nms(np.zeros((0, 7), np.float32), np.zeros((0,), np.float32), 0.5).size == 0
    one = np.array([[1, 2, 0, 4, 2, 1, 0.3]], np.float32)
    assert O.nms(one, np.array([0.9], np.float32), 0.5).tolist() == [0]
    dup = np.repeat(one, 70, 0)  # crosses a 64-box block boundary
    s = np.linspace(1, 0.1, 70).astype(np.float32)
    assert O.nms(dup, s, 0.5).tolist() == [0]
    assert O.nms(dup, s, 1.0).tolist() == list(range(70))  # strict >: IoU == 1.0 does not suppress at thresh 1.0
    assert O.nms(dup, s, 0.5, pre_maxsize=10).tolist() == [0]


def test_flavors_agree_to_conditioning_level():
    """The two reference builds differ by up to a few 1e-5 on small far boxes."""
    a, b = synth.clustered_pairs(120, 120, 9, (35, 17.5), synth.KITTI_PRIORS[:1])
    d = np.abs(O.boxes_iou_bev(a, b, 0) - O.boxes_iou_bev(a, b, 1))
    assert d.max() < 2e-5 and d.max() > 0.0


def test_libdevice_sincos_restatement_accuracy():
    xs = np.concatenate([np.linspace(-7, 7, 3001), np.linspace(-100, 100, 1001), [0.0, np.pi, -np.pi / 2, 1.57]]).astype(np.float32)
    for x in xs:
        s, c = O.sinf(x, 1), O.cosf(x, 1)
        ulp_s = max(abs(np.spacing(np.float32(np.sin(np.float64(x))))), 1e-45)
        ulp_c = max(abs(np.spacing(np.float32(np.cos(np.float64(x))))), 1e-45)
        assert abs(s - np.sin(np.float64(x))) <= 2.0 * ulp_s + 1e-9
        assert abs(c - np.cos(np.float64(x))) <= 2.0 * ulp_c + 1e-9
        # odd / even symmetry is exact, which is what lets cos(-h), sin(-h) be hoisted per box
        assert O.sinf(-x, 1) == -s and O.cosf(-x, 1) == c


def test_points_first_hit_and_margins():
    box = np.array([[[0, 0, 0, 4, 2, 1.5, 0.0], [0, 0, 0, 4, 2, 1.5, 0.0]]], np.float32)  # duplicate: index 0 wins
    pts = np.array([[[0, 0, 0], [2.000005, 0, 0], [2.00002, 0, 0], [0, 0, 0.75], [0, 0, 0.7500001], [2.005, 0, 0]]], np.float32)
    idx = O.points_in_boxes_idx(pts, box, O.FLAVOR_CUDA)[0]
    assert idx.tolist() == [0, 0, -1, 0, -1, -1]  # x open with 1e-5 margin, z closed without margin
    m = O.points_in_boxes_mask(pts[0], box[0], 1e-2, O.FLAVOR_CPU)
    assert m[:, 5].tolist() == [1, 1]  # 1 cm margin of the CPU form


# ---- "next" row 8f-3: RoI-aware pooling / RoI point pooling -------------------------------------------------------
def test_pool_oracle_hand_checked_case():
    """One axis-aligned box, four points whose voxels can be read off by hand (out 2x2x2 over a 4 x 2 x 2 box)."""
    rois = np.array([[0, 0, 0, 4, 2, 2, 0]], np.float32)
    pts = np.array([[-1.5, -0.5, -0.5],   # voxel (0, 0, 0)
                    [1.0, 0.5, 0.5],      # voxel (1, 1, 1)
                    [1.9, 0.9, 0.9],      # voxel (1, 1, 1) again
                    [0.0, 0.0, 1.5],      # outside in z
                    [2.000005, -0.5, -0.5]  # inside only by the 1e-5 margin: (x + dx/2) / res = 2.0000025 -> clamped to voxel 1
                    ], np.float32)
    feat = np.array([[1, 10], [2, 30], [3, 20], [9, 99], [4, 5]], np.float32)
    pooled, argmax, pidx = O.roiaware_pool3d_forward(rois, pts, feat, 2, 3, "max")
    assert pidx[0, 0, 0, 0].tolist() == [1, 0, 0]
    assert pidx[0, 1, 1, 1].tolist() == [2, 1, 2]
    assert pidx[0, 1, 1, 0].tolist() == [0, 0, 0] and pidx[0, 1, 1, 1, 0] == 2
    assert pidx[0, 1, 1, 1, 0] + pidx[0, 0, 0, 0, 0] + pidx[0, 1, 1, 0, 0] + pidx[0, 1, 0, 1, 0] + pidx[0, 1, 0, 0, 0] == int(pidx[..., 0].sum())
    assert pooled[0, 1, 1, 1].tolist() == [3, 30] and argmax[0, 1, 1, 1].tolist() == [2, 1]
    assert argmax[0, 0, 1, 0].tolist() == [-1, -1] and pooled[0, 0, 1, 0].tolist() == [0, 0]
    assert int(pidx[..., 0].sum()) == 4  # the margin point is in, the z-outlier is not
    avg, _, _ = O.roiaware_pool3d_forward(rois, pts, feat, 2, 3, "avg")
    assert avg[0, 1, 1, 1].tolist() == [2.5, 25.0]
    # max_pts 2 keeps one point per voxel: the FIRST in point order (kernel.cu:96-99)
    p2, a2, i2 = O.roiaware_pool3d_forward(rois, pts, feat, 2, 2, "max")
    assert i2[0, 1, 1, 1].tolist() == [1, 1] and p2[0, 1, 1, 1].tolist() == [2, 30]
    g = np.ones_like(pooled)
    gin = O.roiaware_pool3d_backward(pidx, argmax, g, 5, "max")
    assert gin.tolist() == [[1, 1], [0, 1], [1, 0], [0, 0], [1, 1]]
    gin = O.roiaware_pool3d_backward(pidx, argmax, g, 5, "avg")
    assert gin.tolist() == [[1, 1], [0.5, 0.5], [0.5, 0.5], [0, 0], [1, 1]]
    # RoI point pooling: first S inside points in order, cyclic repeat, empty flag
    boxes = np.stack([rois[0], rois[0] + np.array([100, 0, 0, 0, 0, 0, 0], np.float32)])[None]
    pooled, flag = O.roipoint_pool3d_forward(pts[None], feat[None], boxes, 6)
    assert flag.tolist() == [[0, 1]]
    assert pooled[0, 0, :, 3].tolist() == [1, 2, 3, 4, 1, 2] and pooled[0, 0, 1, :3].tolist() == pts[1].tolist()
    assert not pooled[0, 1].any()
    pooled, flag = O.roipoint_pool3d_forward(pts[None], feat[None], boxes, 2)
    assert pooled[0, 0, :, 3].tolist() == [1, 2]


def test_pool_oracle_equals_reference_cuda_golden():
    """The reference has no CPU build of these functions: the restatement is pinned against outputs of the reference
    CUDA kernels on a B200 (tests/golden/make_golden_gpu_pool.py)."""
    import sys

    p = os.path.join(HERE, "golden", "golden_gpu_pool.npz")
    if not os.path.exists(p):
        pytest.skip("golden_gpu_pool.npz not generated yet (tests/golden/make_golden_gpu_pool.py via gpurun)")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_gpu_pool as GP

    g = np.load(p)
    for name, pts, rois, feat, out, mp, s in GP.cases():
        for method in ("max", "avg"):
            pooled, argmax, pidx = O.roiaware_pool3d_forward(rois, pts, feat, out, mp, method)
            assert np.array_equal(pidx, g[f"{name}_pts_idx"]), name
            assert np.array_equal(bits(pooled), bits(g[f"{name}_{method}_pooled"])), (name, method)
            if method == "max":
                assert np.array_equal(argmax, g[f"{name}_argmax"]), name
            go = np.random.default_rng(7).standard_normal(pooled.shape).astype(np.float32)
            gin = O.roiaware_pool3d_backward(pidx, g[f"{name}_argmax"], go, pts.shape[0], method)
            want = g[f"{name}_{method}_grad_in"]  # float atomics: unordered sums
            assert np.abs(gin - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max())), (name, method)
        big = rois.copy()
        big[:, 3:6] += np.float32(0.2)
        pf, fl = O.roipoint_pool3d_forward(pts[None], feat[None], big[None], s)
        assert np.array_equal(fl, g[f"{name}_rp_flag"]) and np.array_equal(bits(pf), bits(g[f"{name}_rp_pooled"])), name
