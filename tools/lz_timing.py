#!/usr/bin/env python
"""Developer tool: per-phase cycle breakdown of nms_lazy_kernel.  Run it here first (it builds
lidardetection_b200/liblidargeom_timing.so with -DLG_LZ_TIMING; the .so travels with the gpurun snapshot), then on the GPU box:
    python tools/lz_timing.py --build ; gpurun -- 'LG_LIB_PATH=lidardetection_b200/liblidargeom_timing.so python tools/lz_timing.py'
Phases are timed by thread 0 of every CTA between the kernel's barriers and summed over CTAs."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    from lidardetection_b200.csrc import build as lbuild

    print(lbuild.build(out=os.path.join(ROOT, "lidardetection_b200", "liblidargeom_timing.so"), extra_flags=["-DLG_LZ_TIMING"]))
    sys.exit(0)

import torch  # noqa: E402

from lidardetection_b200 import _lib, synth  # noqa: E402

P, N = 64, 4096
b, s = synth.cfg2(P, N)
boxes, scores = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
order = scores.sort(1, descending=True)[1].contiguous()
L = _lib.lib()
ws = torch.zeros(L.lg_nms_workspace_bytes_ex(P, N, 0, 0), dtype=torch.uint8, device="cuda")
keep = torch.empty((P, N), dtype=torch.int64, device="cuda")
num = torch.zeros((P,), dtype=torch.int32, device="cuda")
for _ in range(2):
    rc = L.lg_nms_rotated_batched(_lib.ptr(boxes), _lib.ptr(order), None, P, N, 0.01, _lib.ptr(ws), ws.numel(), _lib.ptr(keep), _lib.ptr(num), 0, None)
    assert rc == 0
torch.cuda.synchronize()
off = L.lg_nms_stats_offset(P, N)
st = ws[off:off + 256].view(torch.int64).cpu().tolist()
names = ["alive scan", "dense list + conflicts", "candidates + records", "cull sweeps", "polygon rounds (thread 0)", "wait for last round", "deferred pairs",
         "keep list (+ resolve)", "kill + cluster sync", "prologue: records", "prologue: barrier", "epilogue: keep list"]
tot = sum(st[8:20])
print("pairs cull-tested", st[0], "polygon", st[1], "nonzero", st[2], "kept/frame", float(num.float().mean()))
for n, v in zip(names, st[8:20]):
    print(f"{n:28s} {v / max(tot, 1) * 100:5.1f} %   {v / 1.965e3 / (2 * P):8.1f} us per CTA")
print("sum per CTA", tot / 1.965e3 / (2 * P), "us;  passes per CTA: mean", st[20] / (2 * P), "max", st[21], ";  slowest CTA", st[22] / 1.965e3, "us")
