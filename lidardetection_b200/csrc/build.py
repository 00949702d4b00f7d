#!/usr/bin/env python
"""Build lidardetection_b200/liblidargeom.so: plain nvcc, sm_100a only, no torch / pybind dependency."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "liblidargeom.so")
SOURCES = ["lg_api.cu", "lg_iou.cu", "lg_nms.cu", "lg_points.cu", "lg_pool.cu", "lg_kitti.cu", "lg_select.cu"]
HEADERS = ["lg_common.cuh", "lg_geom.cuh", "lg_strip.cuh", "lg_pib.cuh", os.path.join("..", "..", "include", "lidargeom.h")]
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


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    extra = os.environ.get("LG_EXTRA_NVCC_FLAGS", "").split()  # developer experiments only (e.g. -DLG_LZ_G=4)
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
