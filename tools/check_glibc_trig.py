#!/usr/bin/env python
"""Exhaustive check of lg_trig.cuh's restatement of glibc sinf / cosf (the FL = 0 trigonometry of the *_cpu entry points)
against the host's libm: all 2^32 float bit patterns, through the host build of the device header (tests/host_emu).
~1 minute on 8 cores.  TEST INFRASTRUCTURE (dev tool; tests/test_host_emu.py runs a 2^28-point sweep of the same)."""
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.host_emu import build as ebuild  # noqa: E402

L = C.CDLL(ebuild.build())
L.emu_glibc_sweep.restype = C.c_longlong
L.emu_glibc_sweep.argtypes = [C.c_uint, C.c_uint, C.c_longlong, C.POINTER(C.c_uint)]


def part(i, parts):
    fb = C.c_uint(0)
    n = (1 << 32) // parts
    bad = L.emu_glibc_sweep(i * n, 1, n, C.byref(fb))  # ctypes drops the GIL for the call
    return bad, fb.value


if __name__ == "__main__":
    parts = 64
    with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
        res = list(ex.map(lambda i: part(i, parts), range(parts)))
    bad = sum(r[0] for r in res)
    print(f"glibc sinf/cosf restatement vs libm over all 2^32 floats: {bad} inputs differ", [hex(r[1]) for r in res if r[0]][:8])
    sys.exit(1 if bad else 0)
