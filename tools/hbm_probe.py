#!/usr/bin/env python
"""Developer probe (B200): write-only / read-only / copy HBM bandwidth with torch ops, for context beside
MEASURED_PEAKS.json's copy figure (the IoU matrices and points-in-boxes masks are write-dominated)."""
import json
import torch

n = 1 << 30  # 4 GiB of float32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")


def t(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best * 1e-3


res = {
    "fill_write_GBs": n * 4 / t(lambda: x.zero_()) / 1e9,
    "sum_read_GBs": n * 4 / t(lambda: x.sum()) / 1e9,
    "copy_rw_GBs": 2 * n * 4 / t(lambda: y.copy_(x)) / 1e9,
}
print(json.dumps(res))
