"""CPU-only: the C-ABI library loads, exports every symbol include/lidargeom.h declares, validates its
arguments without touching a GPU, and the Python drop-in modules keep the reference's surface."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

from lidardetection_b200 import _lib, sharded, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lidargeom.h")).read()
    return sorted(set(re.findall(r"LG_API[^;(]*?\b(lg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/lidargeom.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert L.lg_version() == 100


def test_no_torch_or_python_dependency_in_the_library():
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split()[0] for ln in out.splitlines() if ln.strip()]
    assert not any("torch" in n or "python" in n or "c10" in n for n in names), names
    assert any(n.startswith("libcudart") for n in names)


def test_workspace_queries():
    L = _lib.lib()
    assert L.lg_iou_workspace_bytes(0, 0) >= 0
    assert L.lg_iou_workspace_bytes(100, 20) >= 120 * 80
    assert L.lg_iou_workspace_bytes(-1, 5) == 0
    n = L.lg_nms_workspace_bytes(64, 4096)
    assert n >= 64 * 4096 * 80 + 64 * 4096 * 64 * 8
    assert L.lg_nms_workspace_bytes(0, 4096) == 0


def test_argument_validation_without_gpu():
    L = _lib.lib()
    err = lambda: L.lg_last_error_string().decode()
    dummy = C.c_void_p(16)  # never dereferenced: validation happens before any launch
    assert L.lg_boxes_iou_bev(dummy, -1, dummy, 5, dummy, 5, dummy, 1 << 20, 0, None) == -1 and "negative" in err()
    assert L.lg_boxes_iou_bev(None, 4, dummy, 5, dummy, 5, dummy, 1 << 20, 0, None) == -1 and "null" in err()
    assert L.lg_boxes_iou_bev(dummy, 4, dummy, 5, dummy, 4, dummy, 1 << 20, 0, None) == -1 and "ld_out" in err()
    assert L.lg_boxes_iou3d(dummy, 4, dummy, 5, dummy, 5, None, 0, 0, None) == -2
    assert L.lg_boxes_overlap_bev(dummy, 4, dummy, 5, dummy, 5, dummy, 8, 0, None) == -2
    # empty problems are fine and launch nothing
    assert L.lg_boxes_iou_bev(None, 0, None, 7, None, 7, None, 0, 0, None) == 0
    assert L.lg_points_in_boxes(None, None, None, 0, 10, 100, None, 0, 0, None) == 0
    assert L.lg_points_in_boxes(dummy, dummy, dummy, 1, 5000, 100, None, 0, 0, None) == -3
    assert L.lg_nms_rotated_batched(dummy, None, None, 1, 300000, 0.1, dummy, 1 << 30, dummy, dummy, 0, None) == -3
    assert L.lg_nms_rotated_batched(dummy, None, None, 2, 128, 0.1, None, 0, dummy, dummy, 0, None) == -2
    assert L.lg_nms_normal_batched(dummy, None, None, -1, 128, 0.1, dummy, 1 << 20, dummy, dummy, 0, None) == -1
    assert L.lg_nms_rotated_batched(None, None, None, 0, 128, 0.1, None, 0, None, None, 0, None) == 0
    assert L.lg_points_in_boxes_mask(dummy, -2, dummy, 3, dummy, 0.01, 0, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "liblidargeom.so"))
    with pytest.raises(ImportError, match="no CPU"):
        _lib.lib()


def test_python_surface_matches_reference_signatures():
    from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
    from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU

    want = {
        U.boxes_bev_iou_cpu: ["boxes_a", "boxes_b"],
        U.boxes_iou_bev: ["boxes_a", "boxes_b"],
        U.boxes_iou3d_gpu: ["boxes_a", "boxes_b"],
        U.nms_gpu: ["boxes", "scores", "thresh", "pre_maxsize", "kwargs"],
        U.nms_normal_gpu: ["boxes", "scores", "thresh", "kwargs"],
        PU.points_in_boxes_cpu: ["points", "boxes"],
        PU.points_in_boxes_gpu: ["points", "boxes"],
    }
    for fn, names in want.items():
        assert list(inspect.signature(fn).parameters) == names, fn.__name__
    # selected by name through getattr(iou3d_nms_utils, cfg.NMS_TYPE) (model_nms_utils.py:17)
    assert getattr(U, "nms_gpu") is U.nms_gpu and getattr(U, "nms_normal_gpu") is U.nms_normal_gpu


def test_wrappers_reject_bad_inputs_before_any_launch():
    from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
    from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU

    with pytest.raises(AssertionError):
        U.boxes_iou_bev(torch.zeros(3, 6), torch.zeros(2, 7))
    with pytest.raises(AssertionError):
        U.boxes_iou3d_gpu(torch.zeros(3, 7), torch.zeros(2, 7))  # CPU tensors into a GPU function
    with pytest.raises(AssertionError):
        U.nms_gpu(torch.zeros(3, 5), torch.zeros(3), 0.1)
    with pytest.raises(AssertionError):
        PU.points_in_boxes_gpu(torch.zeros(1, 4, 3), torch.zeros(2, 4, 7))
    with pytest.raises(AssertionError):
        PU.points_in_boxes_cpu(np.zeros((4, 2), np.float32), np.zeros((1, 7), np.float32))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65, 200000, 321408):
        for w in (1, 2, 3, 4, 8):
            spans = [sharded.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1 and s0 <= e0
            assert sum(e - s for s, e in spans) == n


def test_synthetic_generators_are_seeded_and_shaped():
    a, gt = synth.cfg1()
    assert a.shape == (321408, 7) and gt.shape == (20, 7) and a.dtype == np.float32
    assert np.array_equal(synth.cfg1()[1], gt)
    b, s = synth.cfg2(n_frames=2, n_boxes=256)
    assert b.shape == (2, 256, 7) and s.shape == (2, 256)
    assert len(np.unique(s[0])) == 256  # distinct scores: the (unstable) torch sort is then deterministic
    p, r = synth.cfg3(n_frames=1, n_points=1000, n_rois=10)
    assert p.shape == (1, 1000, 3) and r.shape == (1, 10, 7)
    x, y = synth.cfg4(n=1000)
    assert x.shape == y.shape == (1000, 7)
    bb, ss = synth.cfg5(n_frames=2, n_classes=3, n_boxes=50)
    assert bb.shape == (2, 3, 50, 7) and ss.shape == (2, 3, 50)


def test_sorting_networks_sort_every_input():
    """0-1 principle over the compare-exchange lists parsed out of lg_geom.cuh"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("check_networks", os.path.join(ROOT, "tools", "check_networks.py"))
    cn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cn)
    nets = cn.networks()
    assert set(nets) >= {"sort8", "sort16"}
    for name, (n, net) in nets.items():
        assert cn.sorts(n, net), name


def test_compact_cell_list_insert_is_order_independent():
    """lg_pib.cuh pib_compact_insert, restated: any insertion order yields the 3-4 smallest ids (+ overflow marker)"""
    import itertools

    def insert(w, k):
        a, b, c, d = w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, w >> 24
        if k < a:
            a, k = k, a
        if k < b:
            b, k = k, b
        if k < c:
            c, k = k, c
        top = 0xFE if d == 0xFE else (k if d == 0xFF else 0xFE)
        return a | (b << 8) | (c << 16) | (top << 24)

    for ids in ([7], [3, 9], [1, 2, 3], [5, 1, 9, 4], [8, 6, 7, 5, 3], [10, 20, 30, 40, 50, 60]):
        want = sorted(ids)
        results = set()
        for perm in itertools.permutations(ids):
            w = 0xFFFFFFFF
            for k in perm:
                w = insert(w, k)
            results.add(w)
        assert len(results) == 1
        w = results.pop()
        got = [(w >> (8 * i)) & 0xff for i in range(4)]
        if len(ids) <= 4:
            assert got == want + [0xFF] * (4 - len(ids))
        else:
            assert got == want[:3] + [0xFE]


def test_cpu_named_entry_points_say_why_they_cannot_run_without_cuda():
    """boxes_bev_iou_cpu / points_in_boxes_cpu are served by the GPU (no CPU fallback by contract): without a usable CUDA
    context -- this container, or a DataLoader worker forked from a CUDA process -- they raise a LidarGeomError that names the
    function and the remedy instead of failing inside the driver ("DataLoader workers")."""
    import numpy as np
    import pytest
    import torch

    from lidardetection_b200 import _lib
    from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
    from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the functions run")
    a = np.array([[0, 0, 0, 4, 2, 1.5, 0.3]], np.float32)
    with pytest.raises(_lib.LidarGeomError, match="boxes_bev_iou_cpu.*no CPU fallback"):
        U.boxes_bev_iou_cpu(a, a)
    with pytest.raises(_lib.LidarGeomError, match="points_in_boxes_cpu.*no CPU fallback"):
        PU.points_in_boxes_cpu(np.zeros((3, 3), np.float32), a)
