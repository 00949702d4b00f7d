#!/usr/bin/env python
"""Compile the UNMODIFIED reference KITTI-eval rotated-IoU kernel into oracle/_ref/kitti_eval/ (test infrastructure only).

The reference (pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:260-291) is a numba.cuda kernel.  numba needs a
CUDA driver even to *compile* an eagerly-typed kernel, and the dev container has none; so this recipe executes the
reference module where it lies under /root/reference with numba's device query answered by a stand-in (compute capability
10.0) and the one `@cuda.jit(<kernel signature>)` deferred, then asks numba for the kernel's PTX (numba -> NVVM of CUDA
12.9) and assembles it with the image's ptxas for sm_100a.  Outputs (git-ignored, they travel to the GPU box):

  oracle/_ref/kitti_eval/rotate_iou_kernel_eval.ptx     what numba/NVVM generate from the reference source
  oracle/_ref/kitti_eval/rotate_iou_kernel_eval.cubin   ptxas -arch=sm_100a of that PTX
  oracle/_ref/kitti_eval/meta.json                      mangled entry name, parameter layout

oracle/ref_kitti.py launches the cubin through the driver API on the GPU box; nothing of the reference is copied into
the repository.
"""
import importlib.util
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("LG_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF_ROOT, "pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py")
OUT = os.path.join(HERE, "_ref", "kitti_eval")
PTXAS = os.environ.get("PTXAS", "/usr/local/cuda/bin/ptxas")


def built():
    return all(os.path.exists(os.path.join(OUT, f)) for f in ("rotate_iou_kernel_eval.cubin", "rotate_iou_kernel_eval.ptx", "meta.json"))


def build(verbose=False):
    if not os.path.exists(SRC):
        print(f"[build_ref_kitti] {SRC} not present; keeping whatever is prebuilt in {OUT}")
        return False
    from numba import cuda
    from numba.core import sigutils
    import numba.cuda.api as api
    import numba.cuda.dispatcher as disp

    class _Dev:  # numba only reads the compute capability while compiling
        compute_capability = (10, 0)
        id = 0

    fake = lambda *a, **k: _Dev()  # noqa: E731
    saved = (api.get_current_device, cuda.get_current_device, cuda.jit, getattr(disp, "get_current_device", None))
    kernels = {}

    def jit(*a, **k):
        if not k.get("device", False) and a and isinstance(a[0], str):  # the __global__ kernel: defer
            def deco(f):
                kernels[f.__name__] = (f, a[0], k)
                return f
            return deco
        return saved[2](*a, **k)

    api.get_current_device = fake
    cuda.get_current_device = fake
    if saved[3] is not None:
        disp.get_current_device = fake
    cuda.jit = jit
    try:
        spec = importlib.util.spec_from_file_location("_ref_rotate_iou", SRC)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        fn, sig, kw = kernels["rotate_iou_kernel_eval"]
        args, _ = sigutils.normalize_signature(sig)
        ptx, _ = cuda.compile_ptx(fn, args, cc=(10, 0), fastmath=bool(kw.get("fastmath", False)))
    finally:
        api.get_current_device, cuda.get_current_device, cuda.jit = saved[:3]
        if saved[3] is not None:
            disp.get_current_device = saved[3]
    os.makedirs(OUT, exist_ok=True)
    ptx_path = os.path.join(OUT, "rotate_iou_kernel_eval.ptx")
    cubin_path = os.path.join(OUT, "rotate_iou_kernel_eval.cubin")
    with open(ptx_path, "w") as f:
        f.write(ptx)
    subprocess.check_call([PTXAS, "-arch=sm_100a", "-O3", ptx_path, "-o", cubin_path] + (["-v"] if verbose else []))
    entry = re.search(r"\.visible \.entry (\w+)\(", ptx).group(1)
    meta = {
        "entry": entry,
        # numba's kernel ABI: scalars as they are, every 1-D array as (meminfo, parent, nitems, itemsize, data, shape0, stride0)
        "params": ["N:i64", "K:i64", "boxes:array", "query_boxes:array", "iou:array", "criterion:i32"],
        "block": 64,
        "source": "pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:260-291",
    }
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"[build_ref_kitti] built {cubin_path}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) else 1)
