"""CPU tier: the DEVICE per-pair code (lidardetection_b200/csrc/lg_geom.cuh, lg_pib.cuh) compiled for the host
(tests/host_emu/, every CUDA intrinsic mapped to the IEEE operation it denotes, libdevice sinf/cosf taken from
the oracle's restatement) must agree with the oracle BIT FOR BIT.  This is what lets the arithmetic contract of
the kernels be iterated on without a GPU; the GPU tier then only has to show that nvcc compiles the same
header to the same arithmetic (tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

from lidardetection_b200 import synth
from oracle import lg_oracle as O
from tests.host_emu import build as ebuild

fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(ebuild.build())
    L.emu_pairs.argtypes = [fp, C.c_int64, fp, C.c_int64, fp, C.c_int, C.c_int]
    L.emu_slow_count.restype = C.c_longlong
    L.emu_points_in_boxes.restype = C.c_longlong
    L.emu_points_in_boxes.argtypes = [fp, C.c_int, fp, C.c_longlong, ip, C.c_int, C.POINTER(C.c_int)]
    L.emu_glibc_sweep.restype = C.c_longlong
    L.emu_glibc_sweep.argtypes = [C.c_uint, C.c_uint, C.c_longlong, C.POINTER(C.c_uint)]
    L.emu_points_mask.argtypes = [fp, C.c_int, fp, C.c_longlong, ip, C.c_int]
    L.emu_trig_symmetry_sweep.restype = C.c_longlong
    L.emu_trig_symmetry_sweep.argtypes = [C.c_uint, C.c_uint, C.c_longlong, C.c_int, C.POINTER(C.c_uint)]
    return L


def pairs(L, a, b, mode, flavor=1):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.empty((len(a), len(b)), np.float32)
    L.emu_pairs(a.ctypes.data_as(fp), len(a), b.ctypes.data_as(fp), len(b), out.ctypes.data_as(fp), mode, flavor)
    return out


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


SETS = {
    "car35": lambda: synth.clustered_pairs(300, 300, 1, (35, 17.5), synth.KITTI_PRIORS[:1]),
    "ped70": lambda: synth.clustered_pairs(300, 300, 2, (70, 35), synth.KITTI_PRIORS[1:2]),
    "mix150": lambda: synth.clustered_pairs(300, 300, 3, (150, 75)),
    "dense": lambda: synth.dense_overlap(300, 300),
    "cfg3iou": synth.cfg3_iou,
    "cfg1sub": lambda: (synth.cfg1()[0][::97], synth.cfg1()[1]),
    "nms": lambda: (synth.cfg2(1, 512)[0][0],) * 2,
}


@pytest.mark.parametrize("name", sorted(SETS))
def test_device_pair_code_matches_oracle_bit_for_bit(emu, name):
    a, b = SETS[name]()
    for mode, ora in ((0, O.boxes_overlap_bev), (1, O.boxes_iou_bev), (2, O.boxes_iou3d)):
        got = pairs(emu, a, b, mode + 4)  # +4: with the exact-zero cull in front, as the kernels run it
        want = ora(a, b, O.FLAVOR_CUDA)
        assert np.array_equal(bits(got), bits(want)), f"{name} mode {mode}: {(bits(got) != bits(want)).sum()} of {got.size} differ"


def test_device_restatement_of_glibc_sinf_cosf_matches_the_host_libm(emu):
    """FL = 0 trigonometry (lg_trig.cuh) against the libm the reference's CPU build calls: every 16th float bit pattern
    (2^28 inputs, each through sinf and cosf; tools/check_glibc_trig.py is the exhaustive sweep), plus every float in the
    range of box headings."""
    fb = C.c_uint(0)
    assert emu.emu_glibc_sweep(7, 16, 1 << 28, C.byref(fb)) == 0, hex(fb.value)
    lo, hi = np.float32(0.78).view(np.uint32), np.float32(6.5).view(np.uint32)  # [0.78, 6.5]: pi/4 .. beyond 2 pi, every float
    for sign in (0, 1 << 31):
        assert emu.emu_glibc_sweep(int(lo) | sign, 1, int(hi - lo), C.byref(fb)) == 0, hex(fb.value)


def test_trig_of_the_negated_heading(emu):
    """make_record takes cosf(-th), sinf(-th) (check_in_box2d, kernel.cu:55-56) as cosf(th), -sinf(th): both trig implementations
    (the oracle's restatement of libdevice, and glibc's restated for the device) are exactly even / odd.  Every 8th positive
    float here; exhaustively: 0 violations for glibc, and only x = +0 for the libdevice restatement (sinf(-0) = +0 there; IEEE
    and the device say -0, which is what the sign flip gives -- the sign of a zero sine never reaches a result)."""
    fb = C.c_uint(0)
    for flavor in (0, 1):
        assert emu.emu_trig_symmetry_sweep(3, 8, 1 << 28, flavor, C.byref(fb)) == 0, (flavor, hex(fb.value))


def test_device_restatement_of_glibc_atan2f_matches_the_host_libm(emu):
    """the literal vertex ordering of the strict (CPU-build) flavor: 4e6 operand pairs -- random bit patterns, coordinate-like
    differences, axis cases -- bit for bit (a 4e8-pair sweep of the same code in C: 0 differences)"""
    r = np.random.default_rng(17)
    n = 2_000_000
    y = np.concatenate([r.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32), r.normal(0, 3, n).astype(np.float32)])
    x = np.concatenate([r.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32), r.normal(0, 3, n).astype(np.float32)])
    y[:64], x[:64] = 0.0, r.normal(0, 1, 64)
    y[64:128], x[64:128] = r.normal(0, 1, 64), 0.0
    x[128:192] = 1.0
    x[192:256] = -x[192:256].__abs__() * 1e-30
    finite = np.isfinite(x) & np.isfinite(y)  # infinities are not special-cased in the device code
    y, x = np.ascontiguousarray(y[finite]), np.ascontiguousarray(x[finite])
    emu.emu_glibc_atan2f_check.restype = C.c_longlong
    emu.emu_glibc_atan2f_check.argtypes = [fp, fp, C.c_longlong]
    assert emu.emu_glibc_atan2f_check(y.ctypes.data_as(fp), x.ctypes.data_as(fp), len(y)) == 0


@pytest.mark.parametrize("name", sorted(SETS))
def test_strict_flavor_matches_the_reference_cpu_build_bit_for_bit(emu, name):
    """LG_FLAG_STRICT_FP32 (boxes_bev_iou_cpu): un-contracted arithmetic + glibc trigonometry == the oracle's CPU flavor, which
    is pinned bit-for-bit to the reference's compiled boxes_iou_bev_cpu (tests/test_oracle_pin.py)"""
    a, b = SETS[name]()
    got = pairs(emu, a, b, 1 + 4, flavor=0)
    want = O.boxes_iou_bev(a, b, O.FLAVOR_CPU)
    assert np.array_equal(bits(got), bits(want)), f"{name}: {(bits(got) != bits(want)).sum()} of {got.size} differ"


def test_strict_points_mask_matches_the_reference_cpu_build(emu):
    pts, rois = synth.cfg3(n_frames=1, n_points=4096, n_rois=60)
    r = np.random.default_rng(3)
    p = np.concatenate([pts[0], boundary_points(r, rois[0], 4096)]).astype(np.float32)
    out = np.empty((60, len(p)), np.int32)
    emu.emu_points_mask(np.ascontiguousarray(rois[0]).ctypes.data_as(fp), 60, p.ctypes.data_as(fp), len(p), out.ctypes.data_as(ip), 0)
    want = O.points_in_boxes_mask(p, rois[0], margin=1e-5, flavor=O.FLAVOR_CPU)
    assert np.array_equal(out, want)


def test_cull_never_drops_a_nonzero_pair(emu):
    r = np.random.default_rng(11)
    a = synth.gt_boxes(400, 5, x_range=(-30, 30), y_range=(-30, 30))
    a[:, 3] = r.uniform(0.3, 10.3, 400)
    a[:, 4] = r.uniform(0.3, 4.3, 400)
    with_cull, without = pairs(emu, a, a, 0 + 4), pairs(emu, a, a, 0)
    assert np.array_equal(bits(with_cull), bits(without))
    assert 0.01 < (without > 0).mean() < 0.2


@pytest.mark.parametrize("jitter", [1e-4, 1e-3, 5e-3, 2e-2])
def test_near_coincident_boxes_take_the_literal_path_and_still_match(emu, jitter):
    """> 8 polygon vertices and angular near-ties are deferred to overlap_area_slow (atan2f + stable order)"""
    r = np.random.default_rng(5)
    base = synth.gt_boxes(200, 11)

    def jit():
        d = base.copy()
        d[:, 0:2] += r.normal(0, jitter, (200, 2))
        d[:, 3:5] *= 1 + r.normal(0, jitter * 0.3, (200, 2))
        d[:, 6] += r.normal(0, jitter, 200)
        return d.astype(np.float32)

    a, b = jit(), jit()
    before = emu.emu_slow_count()
    got = pairs(emu, a, b, 1 + 4)
    assert emu.emu_slow_count() - before >= 10
    assert np.array_equal(bits(got), bits(O.boxes_iou_bev(a, b, O.FLAVOR_CUDA)))


def test_known_answers_through_the_device_code(emu):
    a = np.array([[0, 0, 0, 4, 2, 1.5, 0]], np.float32)
    cases = {(1, 0.5, 0, 4, 2, 1.5, 0.7): 0.414507687, (0, 0, 0, 4, 2, 1.5, np.pi / 2): 0.333333343, (4.005, 0, 0, 4, 2, 1.5, 0): 6.2540517e-4,
             (4.02, 0, 0, 4, 2, 1.5, 0): 0.0, (4, 0, 0, 4, 2, 1.5, 0): 0.0, (50, 50, 0, 4, 2, 1.5, 0): 0.0, (0, 0, 0, 4, 2, 1.5, 0): 1.0}
    for b, want in cases.items():
        got = float(pairs(emu, a, np.array([b], np.float32), 1 + 4, flavor=0)[0, 0])
        assert abs(got - want) < 2e-7, (b, got, want)


# ------------------------------------------------------------------------------------------ points
def pib(L, boxes, pts, grid_words=4096):
    boxes, pts = np.ascontiguousarray(boxes, np.float32), np.ascontiguousarray(pts, np.float32)
    out, used = np.empty(len(pts), np.int32), C.c_int(0)
    tests = L.emu_points_in_boxes(boxes.ctypes.data_as(fp), len(boxes), pts.ctypes.data_as(fp), len(pts), out.ctypes.data_as(ip),
                                  grid_words, C.byref(used))
    return out, tests, used.value


def boundary_points(r, b, m):
    """points inside / on the faces (+- a few ulp) / just outside randomly chosen boxes"""
    k = r.integers(0, len(b), m)
    loc = r.uniform(-0.52, 0.52, (m, 3))
    edge, ax, sgn = r.random(m) < 0.5, r.integers(0, 2, m), r.choice([-1, 1], m)
    loc[edge, 0] = np.where(ax[edge] == 0, sgn[edge] * (0.5 + r.normal(0, 2e-6, edge.sum())), loc[edge, 0])
    loc[edge, 1] = np.where(ax[edge] == 1, sgn[edge] * (0.5 + r.normal(0, 2e-6, edge.sum())), loc[edge, 1])
    loc *= b[k, 3:6]
    c, s = np.cos(b[k, 6]), np.sin(b[k, 6])
    return np.stack([b[k, 0] + loc[:, 0] * c - loc[:, 1] * s, b[k, 1] + loc[:, 0] * s + loc[:, 1] * c, b[k, 2] + loc[:, 2]], 1).astype(np.float32)


@pytest.mark.parametrize("T,M,scale,spread", [(1, 3000, 1, 1), (7, 8000, 0.2, 1), (100, 16384, 1, 1), (254, 10000, 3, 1), (200, 10000, 1, 0.08),
                                               (255, 4000, 1, 1), (1000, 3000, 1, 1)])
def test_grid_cull_of_points_in_boxes_is_exact(emu, T, M, scale, spread):
    """spread = 0.08 packs 200 boxes into a few metres: most cells overflow their four list slots"""
    r = np.random.default_rng(T)
    b = synth.gt_boxes(T, int(r.integers(1 << 30)))
    b[:, 0:2] = b[:, 0:2].mean(0) + (b[:, 0:2] - b[:, 0:2].mean(0)) * spread
    b[:, 3:6] *= scale
    p = boundary_points(r, b, M)
    if T >= 100:  # padded all-zero box, negative size, NaN heading, NaN centre: none may break the cull
        b[3] = 0
        b[5, 3] = -1.0
        b[7, 6] = np.nan
        b[9, 0] = np.nan
    got, tests, used = pib(emu, b, p)
    want = O.points_in_boxes_idx(p[None], b[None], O.FLAVOR_CUDA)[0]
    assert used == (1 if T <= 254 else 0) and np.array_equal(got, want)
    assert tests < 0.3 * M * T or T < 8 or T > 254  # the grid prunes


def test_grid_cull_cfg3_and_fallbacks(emu):
    pts, rois = synth.cfg3(n_frames=2)
    for f in range(2):
        got, tests, used = pib(emu, rois[f], pts[f])
        assert used == 1 and np.array_equal(got, O.points_in_boxes_idx(pts[f:f + 1], rois[f:f + 1], O.FLAVOR_CUDA)[0])
        assert tests < 1.0 * len(got)  # < 1 predicate evaluation per point (the reference: up to 100)
    r = np.random.default_rng(1)
    b = synth.gt_boxes(50, 3)
    b[4, 3] = np.inf  # an unbounded footprint disables the grid: every box is tested
    p = np.concatenate([r.uniform(-10, 80, (3000, 2)), r.uniform(-2, 0, (3000, 1))], 1).astype(np.float32)
    got, _, used = pib(emu, b, p)
    assert used == 0 and np.array_equal(got, O.points_in_boxes_idx(p[None], b[None], O.FLAVOR_CUDA)[0])
    z = np.zeros((5, 7), np.float32)  # only padded boxes: the origin matches box 0 (dataset.py:172-177 padding is not skipped)
    p0 = np.zeros((10, 3), np.float32)
    p0[1:] = r.normal(0, 1e-3, (9, 3))
    got, _, _ = pib(emu, z, p0)
    assert np.array_equal(got, O.points_in_boxes_idx(p0[None], z[None], O.FLAVOR_CUDA)[0]) and got[0] == 0
