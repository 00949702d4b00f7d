#!/usr/bin/env python
"""Join an ncu report's SASS page with nvdisasm line info: executed instructions, lane efficiency and
stall samples per SOURCE LINE of one kernel.  Developer tool (runs in the CPU container).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-name-substring> [--so lib.so] [--top 40]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(so, kernel_sub):
    """offset -> (file, line) for the first function whose mangled name contains kernel_sub"""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, fl, out, fn = None, None, {}, None
        for ln in dis.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if out:
                    return out, fn
                cur = m.group(1) if kernel_sub in m.group(1) else None
                fn = cur
                continue
            if cur is None:
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if m:
                fl = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[int(m.group(1), 16)] = (fl, m.group(2).strip())
        if out:
            return out, fn
    raise SystemExit(f"kernel {kernel_sub} not found in {so}")


def mangled_of(demangled, kernel):
    """void lg::nms_lazy_kernel<(int)1, (bool)1, (int)512>(...) -> nms_lazy_kernelILi1ELb1ELi512EE: the template instance the
    report's launch belongs to (several instances of one kernel share the substring `kernel`)"""
    m = re.search(re.escape(kernel) + r"<([^>]*)>", demangled)
    if not m:
        return kernel
    out = []
    for arg in m.group(1).split(","):
        t = re.fullmatch(r"\s*\((int|bool|unsigned int|long)\)(-?\d+)\s*", arg)
        if not t:
            return kernel
        code = {"int": "i", "bool": "b", "unsigned int": "j", "long": "l"}[t.group(1)]
        v = t.group(2)
        out.append("L" + code + (v if not v.startswith("-") else "n" + v[1:]) + "E")
    return kernel + "I" + "".join(out) + "E"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "lidardetection_b200", "liblidargeom.so"))
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--launch", type=int, default=0, help="which profiled launch of that kernel")
    ap.add_argument("--mangled", default=None, help="substring of the mangled name (default: same as kernel)")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    # the csv holds one block per profiled launch: "Kernel Name",...  then a header row, then rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    blocks = [b for b in blocks if a.kernel in b["name"]]
    # ncu prints every launch twice on this page (same name, same rows): keep one of each pair
    if len(blocks) % 2 == 0 and all(blocks[i]["name"] == blocks[i + 1]["name"] and len(blocks[i]["rows"]) == len(blocks[i + 1]["rows"])
                                    for i in range(0, len(blocks), 2)):
        blocks = blocks[::2]
    b = blocks[a.launch]
    lm, fn = line_map(a.so, a.mangled if a.mangled not in (None, "auto") else mangled_of(b["name"], a.kernel))
    h = b["hdr"]
    ia, ie, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    base = int(b["rows"][0][ia], 16)
    per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_i = tot_t = tot_s = 0
    for r in b["rows"]:
        off = int(r[ia], 16) - base
        fl = lm.get(off, (("?", 0), ""))[0]
        e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
        p = per[fl]
        p[0] += e
        p[1] += t
        p[2] += s
        for i, c in stall_cols:
            v = int(r[i] or 0)
            if v:
                p[3][c[6:]] += v
        tot_i += e
        tot_t += t
        tot_s += s
    print(f"{b['name'][:100]}\n  warp-instr {tot_i}  thread-instr {tot_t}  lanes/instr {tot_t / max(tot_i, 1):.1f}  samples {tot_s}  static instrs {len(lm)}")
    print(f"{'file:line':28s} {'warp-instr':>12s} {'%':>6s} {'lanes':>6s} {'samples':>8s} {'%':>6s}  top stalls")
    for fl, (e, t, s, st) in sorted(per.items(), key=lambda kv: -kv[1][2])[: a.top]:
        tops = " ".join(f"{k}:{v}" for k, v in st.most_common(4))
        print(f"{fl[0] + ':' + str(fl[1]):28s} {e:12d} {100 * e / tot_i:6.2f} {t / max(e, 1):6.1f} {s:8d} {100 * s / max(tot_s, 1):6.2f}  {tops}")


if __name__ == "__main__":
    main()
