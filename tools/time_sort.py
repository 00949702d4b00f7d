import sys, torch
sys.path.insert(0, '/root/repo')
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
for P, N in ((64, 4096), (16, 4096), (256, 1000), (1, 4096)):
    sc = torch.rand(P, N, device='cuda')
    print(P, N, "select_topk %.1f us" % t(lambda: U._argsort_desc(sc)), " torch.sort %.1f us" % t(lambda: sc.sort(1, descending=True)[1]))
