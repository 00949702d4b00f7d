"""Developer tool: lg_select_topk as a full descending sort against torch.sort (CUDA events around 50 back-to-back calls; small
problems are bound by the host's launch rate, ~28 us per call)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidardetection_b200 import _lib
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
def native(sc):
    P, N = sc.shape
    L = _lib.lib()
    order = torch.empty((P, N), dtype=torch.int64, device=sc.device); cnt = torch.empty((P,), dtype=torch.int32, device=sc.device)
    ws = torch.empty(L.lg_select_workspace_bytes(P, N), dtype=torch.uint8, device=sc.device)
    rc = L.lg_select_topk(_lib.ptr(sc), P, N, N, 0.0, 0, None, 0, 0, 1, _lib.ptr(order), _lib.ptr(cnt), None, _lib.ptr(ws), ws.numel(), 0, _lib.stream_ptr(sc.device))
    assert rc == 0
    return order
for P, N in ((64, 4096), (256, 1000), (1024, 1000), (2560, 1000), (2560, 4096)):
    sc = torch.rand(P, N, device='cuda')
    print(P, N, "select_topk %.1f us" % t(lambda: native(sc)), " torch.sort %.1f us" % t(lambda: sc.sort(1, descending=True)[1]))
