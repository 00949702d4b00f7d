#!/usr/bin/env python
"""Turn the ncu reports / bench lines that came back in gpurun_out/ into the committed text summaries under profiles/.

    python tools/make_profiles.py r01 <name>=<report.ncu-rep>:<kernel substring>[:<mangled substring>] ...

For every report: the headline raw metrics of the first matching launch (duration, issue / pipe utilisation, occupancy,
DRAM bytes) and the per-source-line table of tools/ncu_lines.py.  Developer tool (CPU container; needs ncu + nvdisasm).
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def raw_metrics(rep, kernel):
    """metrics of the LONGEST profiled launch whose name contains `kernel` (a call may launch a probe or a skipped twin of the
    same kernel first); returns (metrics, demangled name, ordinal of that launch among the matches)"""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    dur = hdr.index("gpu__time_duration.sum")
    match = [r for r in rows[2:] if kernel in r[kn]]
    if not match:
        raise SystemExit(f"{kernel} not in {rep}")
    best = max(range(len(match)), key=lambda i: float(match[i][dur].replace(",", "")))
    r = match[best]
    return [(k, r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr], r[kn], best


def mangled_of(kname, kernel):
    """iou_strip_kernel<1, 0, 1>(...) -> iou_strip_kernelILi1ELi0ELi1E (integer template arguments only)"""
    import re

    m = re.search(re.escape(kernel) + r"<([^>]*)>", kname)
    if not m:
        return kernel
    args = [a.strip().replace("(int)", "") for a in m.group(1).split(",")]
    if not all(re.fullmatch(r"-?\d+", a) for a in args):
        return kernel
    return kernel + "I" + "".join("Li" + (a if not a.startswith("-") else "n" + a[1:]) + "E" for a in args) + "E"


def traffic(tag, specs):
    """profiles/<tag>_ncu_traffic.json: DRAM bytes per launch of the dominant kernel at the BENCH shape of each workload
    (bench.py copies them into roofline.traffic).  specs: workload=<report.ncu-rep>:<kernel substring>"""
    import json

    out = {}
    for spec in specs:
        wl, rest = spec.split("=", 1)
        rep, kernel = rest.split(":")[:2]
        met, kname, _ = raw_metrics(rep, kernel)
        d = {k: v for k, v, _ in met}
        units = {k: u for k, _, u in met}

        def to_bytes(key):
            v, u = float(d[key].replace(",", "")), units[key].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]

        out[wl] = {"kernel": kname.split("(")[0], "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                   "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1, "usecond": 1, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(units["gpu__time_duration.sum"], 1),
                   "report": os.path.basename(rep), "how": "ncu --set full --clock-control none, one launch at the bench shape (tools/prof_target.py)"}
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2], sys.argv[3:])
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for spec in sys.argv[2:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, kernel = parts[0], parts[1]
        met, kname, ordinal = raw_metrics(rep, kernel)
        mangled = parts[2] if len(parts) > 2 else "auto"  # ncu_lines derives it from the typed name in the source page
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kernel, "--mangled", mangled, "--top", "30",
                                "--launch", str(ordinal)],
                               capture_output=True, text=True).stdout
        out = os.path.join(ROOT, "profiles", f"{tag}_{name}.txt")
        with open(out, "w") as f:
            f.write(f"# {name}: ncu --set full --clock-control none --import-source on (B200, gpurun), report {os.path.basename(rep)}\n")
            f.write(f"# kernel: {kname}\n\n")
            for k, v, u in met:
                f.write(f"{k:90s} {v} {u}\n")
            f.write("\n# per source line (tools/ncu_lines.py): executed warp instructions, active lanes, stall samples\n")
            f.write(lines)
        print("wrote", out)


if __name__ == "__main__":
    main()
