#!/usr/bin/env python
"""Compile oracle/lg_oracle.c -> oracle/liblg_oracle.so (gcc, no contraction, no fast-math)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lg_oracle.c")
OUT = os.path.join(HERE, "liblg_oracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fno-math-errno", "-shared", "-fPIC",
           "-Wall", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
