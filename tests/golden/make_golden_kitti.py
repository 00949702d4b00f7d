#!/usr/bin/env python
"""Generate golden_kitti.npz from the UNMODIFIED reference KITTI-eval kernel (rotate_iou.py:260-291, a numba.cuda kernel
compiled by oracle/build_ref_kitti.py into oracle/_ref/kitti_eval/) running on a B200:

    gpurun -- 'python tests/golden/make_golden_kitti.py'      # writes gpurun_out/golden_kitti.npz
    cp gpurun_out/golden_kitti.npz tests/golden/
    python tests/golden/make_golden_kitti.py --cpu-half       # dev container (needs /root/reference): adds the outputs of the
                                                              # reference's numba-CPU functions (d3_box_overlap_kernel applied
                                                              # to the GPU golden, image_box_overlap) -> tests/golden/golden_kitti_cpu.npz

Inputs are regenerated from seeds by the tests (cases() below); only the reference's outputs are stored.  The reference
ships no tests or vectors for this path, so these files are what pins the oracle's restatement (tests/test_oracle_pin.py).
Both the ptxas-12.9 cubin and the driver-JIT of numba's PTX are run; the file records whether they agree bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidardetection_b200 import synth  # noqa: E402

CRITERIA = (-1, 0, 1, 2)
BEV_COLS = [0, 2, 3, 5, 6]


def hand_made():
    """degenerate / boundary configurations (5-parameter boxes)"""
    a = [0.0, 0.0, 4.0, 2.0, 0.0]
    rows = [
        a,
        [0.0, 0.0, 4.0, 2.0, 0.3],        # rotated; against itself = bit-identical boxes
        [1.0, 0.5, 4.0, 2.0, 0.7],
        [0.0, 0.0, 4.0, 2.0, np.pi / 2],  # 90 degree cross
        [0.3, 0.1, 1.0, 0.5, 0.4],        # contained
        [2.0, 0.0, 4.0, 2.0, 0.0],        # half overlap, collinear edges
        [4.0, 0.0, 4.0, 2.0, 0.0],        # edges touching
        [4.005, 0.0, 4.0, 2.0, 0.0],      # 5 mm gap (no margin in this code: 0)
        [3.0, 2.0, 2.0, 2.0, np.pi / 4],  # corner touch
        [50.0, 50.0, 4.0, 2.0, 1.0],      # far
        [0.0, 0.0, 0.0, 0.0, 0.0],        # zero size
        [0.0, 0.0, 4.0, 0.0, 0.2],        # zero width
        [0.0, 0.0, 4.0, 2.0, 0.3 + 2 * np.pi],
        [0.0, 0.0, 4.0, 2.0, -np.pi],
        [0.0, 0.0, 2.0, 4.0, np.pi / 2 + 0.3],  # same rectangle as row 1, other parametrisation
        [1000.0, -2000.0, 4.0, 2.0, 0.5],
        [1000.5, -2000.2, 3.0, 1.5, -0.9],
        [0.0, 0.0, 4.0, 2.0, 1e-4],
        [0.0, 0.0, 80.0, 60.0, 0.1],      # huge box containing many
    ]
    return np.asarray(rows, dtype=np.float32)


def cases():
    """name -> (boxes (N,5) f32, query_boxes (K,5) f32)"""
    out = {}
    g, d = synth.kitti_eval_frames(40, 811)
    out["frames40"] = (np.concatenate(g)[:, BEV_COLS].astype(np.float32), np.concatenate(d)[:, BEV_COLS].astype(np.float32))
    r = np.random.default_rng(812)
    n = 96
    dense = np.empty((2, n, 5), np.float32)
    dense[..., 0:2] = r.normal(0.0, 0.6, (2, n, 2))
    dense[..., 2] = r.uniform(3.0, 4.5, (2, n))
    dense[..., 3] = r.uniform(1.4, 2.0, (2, n))
    dense[..., 4] = r.uniform(-np.pi, np.pi, (2, n))
    out["dense96"] = (dense[0], dense[1])
    h = hand_made()
    out["hand"] = (h, h)
    # jittered copies of the same objects, 2-decimal ground truth against 4-decimal detections, clustered far from the origin
    g2, d2 = synth.kitti_eval_frames(1, 813, gt_range=(12, 12), fp_range=(0, 0))
    gg = g2[0][:, BEV_COLS]
    rep = np.repeat(gg, 12, 0) + np.concatenate([r.normal(0, 0.02, (len(gg) * 12, 2)), r.normal(0, 0.01, (len(gg) * 12, 2)), r.normal(0, 0.01, (len(gg) * 12, 1))], 1)
    out["near_dup"] = (gg.astype(np.float32), np.round(rep, 4).astype(np.float32))
    return out


def d3_case():
    g, d = synth.kitti_eval_frames(40, 811)
    return np.concatenate(g), np.concatenate(d)


def image_case():
    r = np.random.default_rng(814)
    def boxes(n):
        x1 = r.uniform(0, 1200, n); y1 = r.uniform(0, 370, n)
        return np.stack([x1, y1, x1 + r.uniform(5, 300, n), y1 + r.uniform(5, 200, n)], 1)
    return np.round(boxes(70), 2), np.round(boxes(110), 2)


def main_gpu():
    from oracle import ref_kitti

    out = {}
    same = True
    for name, (b, q) in cases().items():
        for c in CRITERIA:
            r1 = ref_kitti.rotate_iou_gpu_eval(b, q, c, jit=False)
            r2 = ref_kitti.rotate_iou_gpu_eval(b, q, c, jit=True)
            eq = np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
            same &= eq
            out[f"{name}_c{c}"] = r1
            if not eq:
                out[f"{name}_c{c}_jit"] = r2
            print(name, c, r1.shape, "nonzero", int((r1 != 0).sum()), "cubin == driver-JIT:", eq)
    out["cubin_equals_driver_jit"] = np.array([same])
    dst = os.path.join(ROOT, "gpurun_out", "golden_kitti.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst)


def main_cpu_half():
    from oracle import ref_kitti

    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_kitti.npz"))
    k = ref_kitti.d3_box_overlap_kernel()
    assert k is not None, "/root/reference is needed for the numba-CPU half"
    G, D = d3_case()
    out = {}
    for c in CRITERIA:
        rinc = gold["frames40_c2"].copy()
        k(G, D, rinc, c)
        out[f"d3_c{c}"] = rinc
    import importlib.util, types  # noqa: E401
    sys.modules.setdefault("_ref_kitti_eval_pkg", types.ModuleType("_ref_kitti_eval_pkg"))
    mod = sys.modules["_ref_kitti_eval_pkg.eval"] if "_ref_kitti_eval_pkg.eval" in sys.modules else None
    if mod is None:
        ref_kitti.d3_box_overlap_kernel()
    src = os.path.join(os.environ.get("LG_REFERENCE_ROOT", "/root/reference"), "pcdet/datasets/kitti/kitti_object_eval_python/eval.py")
    spec = importlib.util.spec_from_file_location("_ref_kitti_eval_pkg.eval", src)
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    a, b = image_case()
    for c in (-1, 0, 1):
        out[f"image_c{c}"] = ev.image_box_overlap(a, b, c)
    dst = os.path.join(ROOT, "tests", "golden", "golden_kitti_cpu.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    if "--cpu-half" in sys.argv:
        main_cpu_half()
    else:
        main_gpu()
