#!/usr/bin/env python
"""Developer tool: time one bench workload's device step for several builds of the library (LG_LIB_PATH), each in its own process.
    python tools/time_variants.py WORKLOAD lib1.so lib2.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1]
for lib in sys.argv[2:]:
    env = dict(os.environ, LG_LIB_PATH=os.path.join(ROOT, lib) if lib != "default" else "")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--no-cpu-baseline", "--no-secondary", "--steps", "10"], env=env,
                       capture_output=True, text=True)
    try:
        d = json.loads([l for l in r.stdout.split("\n") if l.startswith("{")][-1])
        print(f"{lib:50s} {d['value']:14.4f} {d['unit']:10s} {d['ms_per_step']:.4f} ms/step  roofline {d['roofline']['frac']:.3f}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(lib, "FAILED", e, r.stderr[-400:], flush=True)
