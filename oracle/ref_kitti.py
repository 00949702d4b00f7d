"""Run the compiled UNMODIFIED reference KITTI-eval kernel (oracle/_ref/kitti_eval, built by oracle/build_ref_kitti.py).

GPU box: `rotate_iou_gpu_eval(boxes, query_boxes, criterion)` launches the reference's numba kernel -- from the cubin ptxas
made of numba's PTX, or (jit=True) from the PTX through the driver's own JIT, which is what numba does at run time -- with
the reference's launch geometry (rotate_iou.py:293-330: 64 threads, grid (ceil(N/64), ceil(K/64))).
Dev container: `d3_box_overlap_kernel()` returns the reference's numba-CPU function out of eval.py (needs /root/reference).
TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref", "kitti_eval")
_mods = {}


def available():
    return os.path.exists(os.path.join(_DIR, "rotate_iou_kernel_eval.cubin"))


def _cu():
    L = C.CDLL("libcuda.so.1")
    L.cuModuleLoadData.argtypes = [C.POINTER(C.c_void_p), C.c_void_p]
    L.cuModuleGetFunction.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_char_p]
    L.cuLaunchKernel.argtypes = [C.c_void_p] + [C.c_uint] * 6 + [C.c_uint, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    return L


def _function(jit):
    if jit in _mods:
        return _mods[jit]
    import torch

    torch.cuda.init()
    torch.zeros(1, device="cuda")  # make torch's primary context current on this thread
    meta = json.load(open(os.path.join(_DIR, "meta.json")))
    name = "rotate_iou_kernel_eval." + ("ptx" if jit else "cubin")
    image = open(os.path.join(_DIR, name), "rb").read() + b"\0"
    L = _cu()
    mod, fn = C.c_void_p(), C.c_void_p()
    buf = C.create_string_buffer(image, len(image))
    rc = L.cuModuleLoadData(C.byref(mod), buf)
    assert rc == 0, f"cuModuleLoadData({name}) -> {rc}"
    rc = L.cuModuleGetFunction(C.byref(fn), mod, meta["entry"].encode())
    assert rc == 0, f"cuModuleGetFunction -> {rc}"
    _mods[jit] = (L, fn, mod, buf)
    return _mods[jit]


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, jit=False):
    """numpy (N,5), (K,5) -> (N,K) float32, exactly as the reference wrapper (rotate_iou.py:293-330)"""
    import torch

    boxes = np.ascontiguousarray(np.asarray(boxes).astype(np.float32))
    query_boxes = np.ascontiguousarray(np.asarray(query_boxes).astype(np.float32))
    N, K = boxes.shape[0], query_boxes.shape[0]
    iou = np.zeros((N, K), dtype=np.float32)
    if N == 0 or K == 0:
        return iou
    L, fn, _, _ = _function(jit)
    tb = torch.from_numpy(boxes.reshape(-1)).cuda()
    tq = torch.from_numpy(query_boxes.reshape(-1)).cuda()
    to = torch.from_numpy(iou.reshape(-1)).cuda()
    torch.cuda.synchronize()
    vals = [C.c_int64(N), C.c_int64(K)]
    for t in (tb, tq, to):  # numba array ABI: meminfo, parent, nitems, itemsize, data, shape[0], strides[0]
        vals += [C.c_void_p(0), C.c_void_p(0), C.c_int64(t.numel()), C.c_int64(4), C.c_void_p(t.data_ptr()), C.c_int64(t.numel()), C.c_int64(4)]
    vals.append(C.c_int32(int(criterion)))
    params = (C.c_void_p * len(vals))(*[C.cast(C.byref(v), C.c_void_p) for v in vals])
    rc = L.cuLaunchKernel(fn, (N + 63) // 64, (K + 63) // 64, 1, 64, 1, 1, 0, None, params, None)
    assert rc == 0, f"cuLaunchKernel -> {rc}"
    torch.cuda.synchronize()
    return to.cpu().numpy().reshape(N, K)


def d3_box_overlap_kernel():
    """the reference's numba CPU function (eval.py:116-147), loaded without importing the pcdet package"""
    import importlib.util
    import sys
    import types

    root = os.environ.get("LG_REFERENCE_ROOT", "/root/reference")
    src = os.path.join(root, "pcdet/datasets/kitti/kitti_object_eval_python/eval.py")
    if not os.path.exists(src):
        return None
    pkg = types.ModuleType("_ref_kitti_eval_pkg")
    pkg.__path__ = []
    stub = types.ModuleType("_ref_kitti_eval_pkg.rotate_iou")
    stub.rotate_iou_gpu_eval = None  # the CUDA half is run through the cubin above
    sys.modules["_ref_kitti_eval_pkg"] = pkg
    sys.modules["_ref_kitti_eval_pkg.rotate_iou"] = stub
    spec = importlib.util.spec_from_file_location("_ref_kitti_eval_pkg.eval", src)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.d3_box_overlap_kernel
