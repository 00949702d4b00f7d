#!/usr/bin/env python
"""Build lidardetection_b200/liblidargeom.so: plain nvcc, sm_100a only, no torch / pybind dependency."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "liblidargeom.so")
SOURCES = ["lg_api.cu", "lg_iou.cu", "lg_nms.cu", "lg_points.cu", "lg_pool.cu", "lg_kitti.cu", "lg_select.cu"]
HEADERS = ["lg_common.cuh", "lg_geom.cuh", "lg_trig.cuh", "lg_strip.cuh", "lg_pib.cuh", os.path.join("..", "..", "include", "lidargeom.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "--shared", "-cudart", "shared",
]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: developer builds next to the product library (e.g. -DLG_LZ_TIMING into liblidargeom_timing.so)"""
    if out is None and not force and not stale():
        return OUT
    extra = os.environ.get("LG_EXTRA_NVCC_FLAGS", "").split()  # developer experiments only (e.g. -DLG_LZ_G=4)
    # one nvcc per translation unit, side by side (the units are independent: no relocatable device code), then one link
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(PKG, "build" if out is None else "build_" + os.path.splitext(os.path.basename(out))[0])
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in FLAGS if f != "--shared"] + extra + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        subprocess.check_call([NVCC] + cflags + ["-c", "-o", obj, os.path.join(HERE, src)])
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([NVCC] + FLAGS + ["-o", out or OUT] + objs)
    return out or OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
