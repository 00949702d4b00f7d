"""SURVEY 8f-2 ("next" row): the KITTI evaluation's rotated overlaps (kitti_object_eval_python/rotate_iou.py, eval.py:80-155,
340-414).

CPU tier (no GPU): the oracle restatement is PINNED bit for bit against
  * tests/golden/golden_kitti.npz      -- outputs of the unmodified reference numba.cuda kernel run on a B200
                                          (tests/golden/make_golden_kitti.py; cubin of ptxas 12.9 == driver JIT of the PTX),
  * tests/golden/golden_kitti_cpu.npz  -- outputs of the reference's numba CPU functions (d3_box_overlap_kernel applied to
                                          the GPU golden, image_box_overlap), generated in the dev container,
plus hand-derivable known answers, the host-side part logic and the C-ABI argument validation.
GPU tier (`-m gpu`): the product (Python drop-in -> ctypes -> C ABI -> CUDA) against the oracle, the golden vectors and -- when
oracle/_ref travelled -- the reference kernel run live on the same GPU.  Bar: bit-exact (tolerance stated: 0 ulp; the spec's
IoU bar is 1e-5 absolute) for every pair whose polygon has <= 8 points; pairs with more overflow the reference's 8-point
buffer (undefined behaviour there), for those the product must equal the oracle's 24-point restatement.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from lidardetection_b200 import _lib, synth
from lidardetection_b200.datasets.kitti.kitti_object_eval_python import eval as E
from lidardetection_b200.datasets.kitti.kitti_object_eval_python import rotate_iou as R
from oracle import lg_oracle as O
from oracle import ref_kitti

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_kitti as MG  # noqa: E402


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def same_bits(a, b, where=None):
    ne = (bits(a) != bits(b)) & ~(np.isnan(a) & np.isnan(b))
    if where is not None:
        ne &= where
    return int(ne.sum())


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_kitti.npz")), np.load(os.path.join(golden_dir, "golden_kitti_cpu.npz"))


# ------------------------------------------------------------------ CPU tier
@pytest.mark.parametrize("name", ["frames40", "dense96", "hand", "near_dup"])
def test_oracle_bit_exact_vs_reference_numba_kernel(gold, name):
    g, _ = gold
    assert bool(g["cubin_equals_driver_jit"][0])
    b, q = MG.cases()[name]
    cnt = O.rotate_iou_eval_cnt(b, q)
    for c in MG.CRITERIA:
        ref = g[f"{name}_c{c}"]
        mine = O.rotate_iou_eval(b, q, c)
        assert mine.shape == ref.shape
        assert same_bits(mine, ref, cnt <= 8) == 0, (name, c)
    # the un-contracted flavor is the same algorithm: conditioning-level agreement only
    d = np.abs(O.rotate_iou_eval(b, q, 2, O.FLAVOR_CPU).astype(np.float64) - g[f"{name}_c2"])
    assert np.nanmax(np.where(cnt <= 8, d, 0)) < (2e-3 if name != "hand" else np.inf)


def test_oracle_d3_and_image_overlap_vs_reference_numba_cpu(gold):
    _, gc = gold
    G, D = MG.d3_case()
    for c in MG.CRITERIA:
        assert same_bits(O.d3_box_overlap(G, D, c), gc[f"d3_c{c}"]) == 0, c
    a, b = MG.image_case()
    for c in (-1, 0, 1):
        want = gc[f"image_c{c}"]
        got = E.image_box_overlap(a, b, c)
        assert got.dtype == want.dtype and np.array_equal(got, want), c


def test_known_answers():
    a = np.array([[0, 0, 4, 2, 0]], np.float32)
    f = lambda q, c=-1: float(O.rotate_iou_eval(a, np.array([q], np.float32), c)[0, 0])  # noqa: E731
    assert f([0, 0, 4, 2, 0]) == 1.0                       # identical, axis aligned: exact corners
    assert f([2, 0, 4, 2, 0]) == pytest.approx(1 / 3, abs=1e-6)
    assert f([2, 0, 4, 2, 0], 2) == 4.0                    # intersection area
    assert f([2, 0, 2, 2, 0], 0) == pytest.approx(0.5, abs=1e-6)   # criterion 0 divides by the QUERY box's area (4)
    assert f([2, 0, 2, 2, 0], 1) == pytest.approx(0.25, abs=1e-6)  # criterion 1 by the box's (8)
    assert f([0, 0, 4, 2, np.pi / 2]) == pytest.approx(1 / 3, abs=1e-6)
    assert f([4.005, 0, 4, 2, 0]) == 0.0                   # no margin in this code (iou3d_nms has 1e-2)
    assert f([50, 50, 4, 2, 1.0]) == 0.0
    # clockwise-positive angle: a long box turned by +30 degrees reaches the point (1.5, -0.75) but not (1.5, +0.75)
    long_box = np.array([[0, 0, 4, 0.5, np.pi / 6]], np.float32)
    hit = O.rotate_iou_eval(long_box, np.array([[1.5, -0.75, 0.2, 0.2, 0]], np.float32), 2)[0, 0]
    miss = O.rotate_iou_eval(long_box, np.array([[1.5, 0.75, 0.2, 0.2, 0]], np.float32), 2)[0, 0]
    assert hit > 0.03 and miss == 0.0
    # the zero-size-quadrilateral quirk of point_in_quadrilateral (0 >= 0 >= 0): a zero box "contains" every point
    z = O.rotate_iou_eval(np.array([[50, 50, 4, 2, 1.0]], np.float32), np.array([[0, 0, 0, 0, 0]], np.float32), 2)[0, 0]
    assert z == pytest.approx(8.0, abs=1e-4)
    assert O.rotate_iou_eval(np.zeros((0, 5), np.float32), a).shape == (0, 1)


def test_d3_semantics():
    # camera boxes (x, y, z, l, h, w, ry): y is the BOTTOM face, the box spans [y - h, y]
    b = np.array([[0, 2.0, 10, 4, 1.5, 2, 0]], np.float64)
    q = np.array([[0, 1.5, 10, 4, 1.5, 2, 0], [0, 0.4, 10, 4, 1.5, 2, 0], [30, 2.0, 10, 4, 1.5, 2, 0]], np.float64)
    r = O.d3_box_overlap(b, q, -1)
    assert r[0, 0] == pytest.approx(8.0 / (12 + 12 - 8.0), abs=1e-6)  # heights overlap by 1.0
    assert r[0, 1] == 0.0 and r[0, 2] == 0.0                         # no height overlap / no BEV overlap
    assert O.d3_box_overlap(b, q, 2)[0, 0] == 1.0                    # criterion "else": inc / inc


def test_split_parts_and_host_logic():
    assert E.get_split_parts(3769, 50) == [75] * 50 + [19]
    assert E.get_split_parts(100, 50) == [2] * 50
    assert E.get_split_parts(7, 50) == [7]
    assert E.image_box_overlap(np.zeros((0, 4)), np.zeros((3, 4))).shape == (0, 3)


def test_c_abi_validation_without_gpu():
    L = _lib.lib()
    err = lambda: L.lg_last_error_string().decode()  # noqa: E731
    d = C.c_void_p(16)
    assert L.lg_kitti_workspace_bytes(10, 20, 0) >= 30 * 48 and L.lg_kitti_workspace_bytes(-1, 2, 0) == 0
    assert L.lg_rotate_iou_eval(d, -1, d, 3, d, -1, d, 1 << 20, 0, None) == -1 and "negative" in err()
    assert L.lg_rotate_iou_eval(None, 2, d, 3, d, -1, d, 1 << 20, 0, None) == -1 and "null" in err()
    assert L.lg_rotate_iou_eval(d, 2, d, 3, d, -1, None, 0, 0, None) == -2
    assert L.lg_rotate_iou_eval(None, 0, None, 3, None, -1, None, 0, 0, None) == 0
    assert L.lg_d3_box_overlap(d, 2, d, 3, d, -1, d, 8, 0, None) == -2
    assert L.lg_d3_box_overlap(None, 5, None, 0, None, -1, None, 0, 0, None) == 0
    assert L.lg_kitti_overlaps_parts(d, 4, d, 4, d, d, d, 2, 8, 0, -1, d, d, 1 << 20, 0, None) == -1 and "metric" in err()
    assert L.lg_kitti_overlaps_parts(d, 4, d, 4, d, d, d, 2, 0, 1, -1, d, d, 1 << 20, 0, None) == 0


# ------------------------------------------------------------------ GPU tier
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", ["frames40", "dense96", "hand", "near_dup"])
def test_rotate_iou_gpu_eval_vs_oracle_and_golden(gold, name):
    g, _ = gold
    b, q = MG.cases()[name]
    cnt = O.rotate_iou_eval_cnt(b, q)
    for c in MG.CRITERIA:
        ours = R.rotate_iou_gpu_eval(b.astype(np.float64), q.astype(np.float64), c)  # float64 in, like the evaluation's annos
        assert ours.dtype == np.float32 and ours.shape == (len(b), len(q))
        assert same_bits(ours, O.rotate_iou_eval(b, q, c)) == 0, (name, c)         # incl. the > 8-point pairs
        assert same_bits(ours, g[f"{name}_c{c}"], cnt <= 8) == 0, (name, c)
        if ref_kitti.available():
            assert same_bits(ours, ref_kitti.rotate_iou_gpu_eval(b, q, c), cnt <= 8) == 0, (name, c)


@gpu
def test_strict_flag_is_the_uncontracted_flavor():
    import torch

    b, q = MG.cases()["dense96"]
    tb, tq = torch.from_numpy(b).cuda(), torch.from_numpy(q).cuda()
    got = R.rotate_iou_eval_cuda(tb, tq, -1, flags=_lib.LG_FLAG_STRICT_FP32).cpu().numpy()
    want = O.rotate_iou_eval(b, q, -1, O.FLAVOR_CPU)
    # same statements without contraction; cos/sin stay libdevice's on the device -> not bit-comparable, 1e-5 is the bar
    assert np.abs(got - want).max() <= 1e-5


@gpu
def test_d3_box_overlap_vs_oracle_and_reference_cpu_golden(gold):
    _, gc = gold
    G, D = MG.d3_case()
    for c in MG.CRITERIA:
        ours = E.d3_box_overlap(G, D, c)
        assert same_bits(ours, O.d3_box_overlap(G, D, c)) == 0, c
        assert same_bits(ours, gc[f"d3_c{c}"]) == 0, c
    assert same_bits(E.bev_box_overlap(G[:, MG.BEV_COLS], D[:, MG.BEV_COLS]), gold[0]["frames40_c-1"]) == 0


def _annos(frames):
    return [{"name": np.array(["Car"] * len(f)), "location": f[:, 0:3], "dimensions": f[:, 3:6], "rotation_y": f[:, 6],
             "bbox": np.abs(f[:, [0, 1, 0, 1]]) * 10 + np.array([0, 0, 30, 20])} for f in frames]


@gpu
@pytest.mark.parametrize("metric", [0, 1, 2])
def test_calculate_iou_partly_equals_per_part_calls(metric):
    gts, dts = synth.kitti_eval_frames(23, 77)
    gts[5] = gts[5][:0]  # a frame without ground truth, one without detections
    dts[11] = dts[11][:0]
    ga, da = _annos(gts), _annos(dts)
    overlaps, parted, ng, nd = E.calculate_iou_partly(ga, da, metric, num_parts=5)
    assert len(parted) == 6 and len(overlaps) == 23  # 23 = 5 * 4 + 3
    assert list(ng) == [len(x) for x in gts] and list(nd) == [len(x) for x in dts]
    idx = 0
    for p, num in enumerate(E.get_split_parts(23, 5)):
        G, D = np.concatenate(gts[idx:idx + num]), np.concatenate(dts[idx:idx + num])
        if metric == 0:
            want = E.image_box_overlap(np.concatenate([a["bbox"] for a in ga[idx:idx + num]]), np.concatenate([a["bbox"] for a in da[idx:idx + num]]))
        elif metric == 1:
            want = O.bev_box_overlap(G[:, MG.BEV_COLS], D[:, MG.BEV_COLS]).astype(np.float64)
        else:
            want = O.d3_box_overlap(G, D).astype(np.float64)
        assert parted[p].dtype == np.float64 and np.array_equal(parted[p], want, equal_nan=True), (metric, p)
        gi = di = 0
        for i in range(num):
            o = overlaps[idx + i]
            assert o.shape == (len(gts[idx + i]), len(dts[idx + i]))
            assert np.array_equal(o, want[gi:gi + o.shape[0], di:di + o.shape[1]], equal_nan=True)
            gi += o.shape[0]
            di += o.shape[1]
        idx += num


@gpu
def test_many_ragged_parts_cover_every_tile():
    """parts of 0..70 x 0..600 boxes (tiles of 32 x 256: partial, empty and multi-tile parts) against the oracle per part"""
    import torch

    r = np.random.default_rng(3)
    P = 120
    gc = r.integers(0, 71, P)
    dc = r.integers(0, 601, P)
    gc[:6], dc[:6] = [0, 5, 1, 33, 32, 64], [7, 0, 1, 257, 256, 512]
    gts, dts = synth.kitti_eval_frames(1, 9, gt_range=(int(gc.sum()), int(gc.sum())), fp_range=(int(dc.sum()), int(dc.sum())))
    G, D = gts[0], dts[0][:int(dc.sum())]
    D[:, 0:3] = G[r.integers(0, len(G), len(D)), 0:3] + r.normal(0, 1.5, (len(D), 3))  # make overlaps frequent
    for metric in (1, 2):
        out, off = E.kitti_overlaps_parts_cuda(torch.from_numpy(G).cuda(), torch.from_numpy(D).cuda(), gc, dc, metric)
        out = out.cpu().numpy()
        g0 = d0 = 0
        for p in range(P):
            Gp, Dp = G[g0:g0 + gc[p]], D[d0:d0 + dc[p]]
            want = O.bev_box_overlap(Gp[:, MG.BEV_COLS], Dp[:, MG.BEV_COLS]) if metric == 1 else O.d3_box_overlap(Gp, Dp)
            got = out[off[p]:off[p + 1]].reshape(gc[p], dc[p])
            assert same_bits(got, want) == 0, (metric, p, gc[p], dc[p])
            g0 += gc[p]
            d0 += dc[p]
        assert off[-1] == (gc * dc).sum()


@gpu
def test_full_size_eval_properties():
    """KITTI val size (3769 frames, 51 parts, 28 M pairs): properties that need no oracle pass over everything --
    criterion identities between the four outputs of the same pairs, symmetry of the intersection area under swapping the
    roles (up to the reference's own argument-order rounding), and the oracle on a random sample of pairs."""
    import torch

    gts, dts = synth.kitti_eval_frames(3769, 5)
    parts = E.get_split_parts(3769, 50)
    gc, dc, i = [], [], 0
    for n in parts:
        gc.append(sum(len(x) for x in gts[i:i + n]))
        dc.append(sum(len(x) for x in dts[i:i + n]))
        i += n
    G, D = np.concatenate(gts), np.concatenate(dts)
    tg, td = torch.from_numpy(G).cuda(), torch.from_numpy(D).cuda()
    outs = {c: E.kitti_overlaps_parts_cuda(tg, td, gc, dc, 1, c)[0] for c in MG.CRITERIA}
    off = E.kitti_overlaps_parts_cuda(tg, td, gc, dc, 1, 2)[1]
    assert outs[2].numel() == int(off[-1]) > 2.5e7
    inter = outs[2].double()
    assert bool((inter >= 0).all()) and 0.001 < float((inter > 0).double().mean()) < 0.02
    # sample pairs (all the overlapping ones of three parts + random ones) against the oracle
    r = np.random.default_rng(1)
    for p in (0, 17, 50):
        g0, d0 = sum(gc[:p]), sum(dc[:p])
        Gp, Dp = G[g0:g0 + gc[p]][:, MG.BEV_COLS], D[d0:d0 + dc[p]][:, MG.BEV_COLS]
        sel = r.choice(gc[p], 40, replace=False)
        for c in MG.CRITERIA:
            got = outs[c][off[p]:off[p + 1]].reshape(gc[p], dc[p]).cpu().numpy()[sel]
            assert same_bits(got, O.rotate_iou_eval(Gp[sel], Dp, c)) == 0, (p, c)
    # 3-D: never larger than ... the BEV intersection scaled by the height overlap; zero where BEV is zero
    o3 = E.kitti_overlaps_parts_cuda(tg, td, gc, dc, 2, -1)[0]
    assert bool(((o3 > 0) <= (outs[2] > 0)).all()) and float(o3.max()) <= 1.0 + 1e-6
