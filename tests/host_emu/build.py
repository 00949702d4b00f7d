"""Build tests/host_emu/libemu_geom.so: the device per-pair headers compiled for the host (TEST INFRASTRUCTURE)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libemu_geom.so")
SRCS = [os.path.join(HERE, "emu_geom.cpp")]
DEPS = SRCS + [os.path.join(HERE, "lg_host_emu.h"), os.path.join(ROOT, "lidardetection_b200", "csrc", "lg_geom.cuh"),
               os.path.join(ROOT, "lidardetection_b200", "csrc", "lg_pib.cuh"),
               os.path.join(ROOT, "lidardetection_b200", "csrc", "lg_trig.cuh")]


def build(force=False):
    from oracle import build as obuild

    ora = obuild.build()
    deps = [d for d in DEPS if os.path.exists(d)] + [ora]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-shared", "-fPIC", "-I", cuda_inc, "-o", OUT] + SRCS + \
          [ora, "-Wl,-rpath," + os.path.dirname(ora)]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    import sys

    sys.path.insert(0, ROOT)
    print(build(force=True))
