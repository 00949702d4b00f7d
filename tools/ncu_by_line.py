#!/usr/bin/env python
"""Developer tool: aggregate an `ncu --import-source on` report per CUDA source line.
    python tools/ncu_by_line.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top]
The SASS page of the report (per-instruction executed counts and stall samples) is joined, by instruction offset, with the
line table of the same kernel in the local build (nvdisasm -g of the cubin inside liblidargeom.so), so the .so that ran on the
GPU box must be the one in the tree."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre, "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.split("\n")
kname = lines[0].split('","')[1].rstrip('",') if lines and lines[0].startswith('"Kernel Name"') else "?"
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
h = rows[0]
ia, ie, isamp, ithr = h.index("Address"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Thread Instructions Executed")
sass = [(int(r[ia], 16), r[1].strip(), int(r[ie] or 0), int(r[isamp] or 0), int(r[ithr] or 0)) for r in rows[1:] if len(r) > ithr and r[ia].startswith("0x")]
base = sass[0][0]
seen = {}
for x in sass:
    seen.setdefault(x[0], x)
sass = list(seen.values())  # some reports list every instruction twice
nsass = len(sass)
# line table of the local build
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lidardetection_b200", "liblidargeom.so")], cwd=tmp, capture_output=True)
short = re.sub(r"^void ", "", kname).split("(")[0]
mangled_hint = short.split("<")[0].split("::")[-1]
table = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin") or f.count("-") > 0:
        continue
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    secs = [i for i, l in enumerate(txt) if l.startswith("//--------------------- .text.") and mangled_hint in l]
    for si in secs:  # the template instantiation with exactly the report's instruction count
        cur, t = None, {}
        for l in txt[si + 1:]:
            if l.startswith("//--------------------- "):
                break
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+[A-Z@!]", l)
            if m:
                t[int(m.group(1), 16)] = cur or ("?", 0)
        if len(t) == nsass:
            table = t
            break
    if table:
        break
tot_e = sum(x[2] for x in sass)
tot_s = sum(x[3] for x in sass)
by = collections.defaultdict(lambda: [0, 0, 0, 0])
for addr, text, ex, samp, thr in sass:
    k = table.get(addr - base, ("?", 0))
    by[k][0] += ex
    by[k][1] += samp
    by[k][2] += thr
    by[k][3] += 1
print(f"{kname[:100]}\n{len(sass)} SASS instructions, {tot_e} warp-instructions executed, {tot_s} stall samples; line table entries {len(table)}")
print(f"{'file:line':28s} {'exec %':>7s} {'samples %':>9s} {'lanes':>6s} {'#sass':>6s}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    print(f"{k[0] + ':' + str(k[1]):28s} {100.0 * v[0] / max(tot_e, 1):7.2f} {100.0 * v[1] / max(tot_s, 1):9.2f} {v[2] / max(v[0], 1):6.1f} {v[3]:6d}")
# opcode histogram
op = collections.Counter()
for addr, text, ex, samp, thr in sass:
    t = text.split()
    o = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op[o.split(".")[0]] += ex
print("opcodes:", ", ".join(f"{k} {100.0 * v / max(tot_e, 1):.1f}%" for k, v in op.most_common(28)))
