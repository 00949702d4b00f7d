#!/usr/bin/env python
"""Build the UNMODIFIED reference extensions into oracle/_ref/ (test infrastructure only).

This compiles the reference's own source files *where they lie* under /root/reference
(nothing is copied into the repository) with a recipe of ours -- NOT the reference's
setup.py -- and writes only into the git-ignored directory oracle/_ref/:

  oracle/_ref/iou3d_nms_cuda/iou3d_nms_cuda.so          (pcdet/ops/iou3d_nms/src/*.{cpp,cu})
  oracle/_ref/roiaware_pool3d_cuda/roiaware_pool3d_cuda.so (pcdet/ops/roiaware_pool3d/src/*.{cpp,cu})
  oracle/_ref/roipoint_pool3d_cuda/roipoint_pool3d_cuda.so (pcdet/ops/roipoint_pool3d/src/*.{cpp,cu}; "next" row 8f-3)

They are the reference's pybind11 torch extensions, compiled for sm_100a, and serve as
  * tier B oracle (CPU functions boxes_iou_bev_cpu / points_in_boxes_cpu, runnable without GPU),
  * tier A oracle (the reference CUDA kernels run on the same B200 as our kernels),
  * the `cpu_baseline.kind == "reference"` timing arm of bench.py.

`-O2` is mandatory: at -O0 the CPU IoU binds to an nvcc host stub of the same mangled name
(check_rect_cross) whose body is exit(1).

/root/reference does not exist on the GPU box: there, the prebuilt .so files (which travel with
the gpurun snapshot) are simply imported by oracle/ref_loader.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("LG_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "iou3d_nms_cuda": [
        "pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu",
    ],
    "roiaware_pool3d_cuda": [
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp",
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu",
    ],
    "roipoint_pool3d_cuda": [
        "pcdet/ops/roipoint_pool3d/src/roipoint_pool3d.cpp",
        "pcdet/ops/roipoint_pool3d/src/roipoint_pool3d_kernel.cu",
    ],
}


# the reference's own Python wrappers of the two extensions, byte-compiled (py_compile) into oracle/_ref/py/*.pyc so that the
# route-B test (tests/test_route_b.py) can run the UNMODIFIED reference Python on the GPU box, where /root/reference does
# not exist; compiled outputs only, no source is copied
PY_MODULES = {
    "iou3d_nms_utils": "pcdet/ops/iou3d_nms/iou3d_nms_utils.py",
    "roiaware_pool3d_utils": "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py",
}


def build_py():
    import py_compile

    out = os.path.join(OUT, "py")
    os.makedirs(out, exist_ok=True)
    for name, rel in PY_MODULES.items():
        py_compile.compile(os.path.join(REF_ROOT, rel), cfile=os.path.join(out, name + ".pyc.bin"), dfile=rel, doraise=True)
    print(f"[build_ref] byte-compiled {sorted(PY_MODULES)} into {out}")


def py_built():
    return all(os.path.exists(os.path.join(OUT, "py", n + ".pyc.bin")) for n in PY_MODULES)


def build(verbose=False):
    if not os.path.isdir(REF_ROOT):
        print(f"[build_ref] {REF_ROOT} not present; keeping whatever is prebuilt in {OUT}")
        return False
    build_py()
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    ok = True
    for name, rel in EXTS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        srcs = [os.path.join(REF_ROOT, r) for r in rel]
        try:
            load(
                name=name,
                sources=srcs,
                extra_cflags=["-O2", "-w"],
                extra_cuda_cflags=["-O3", "-w", "-lineinfo"],
                build_directory=bdir,
                verbose=verbose,
            )
            print(f"[build_ref] built {bdir}/{name}.so")
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[build_ref] FAILED {name}: {e}", file=sys.stderr)
    return ok


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) else 1)
