import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
wl = bench.PostProcWorkload(torch, 0)
from lidardetection_b200 import model_nms_utils as M
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n
print("whole step", t(wl.step))
names = [a for a in dir(wl) if not a.startswith('_')]
print(names)
sc = [getattr(wl, a) for a in names if isinstance(getattr(wl, a), torch.Tensor) and getattr(wl, a).is_cuda]
print([(x.shape, x.dtype) for x in sc])
scores = [x for x in sc if x.dim() == 2][0]; boxes = [x for x in sc if x.dim() == 3][0]
masked = torch.where(scores >= 0.1, scores, scores.new_full((), float('-inf')))
print("where", t(lambda: torch.where(scores >= 0.1, scores, scores.new_full((), float('-inf')))))
print("topk", t(lambda: torch.topk(masked, k=4096, dim=1)))
ts, ti = torch.topk(masked, k=4096, dim=1)
print("counts", t(lambda: (ts > float('-inf')).sum(1).to(torch.int32)))
print("gather", t(lambda: torch.gather(boxes[:, :, 0:7], 1, ti.unsqueeze(-1).expand(64, 4096, 7)).contiguous()))
tb = torch.gather(boxes[:, :, 0:7], 1, ti.unsqueeze(-1).expand(64, 4096, 7)).contiguous()
cnt = (ts > float('-inf')).sum(1).to(torch.int32)
print("nms", t(lambda: U._nms_call('lg_nms_rotated_batched', tb.float(), None, cnt.contiguous(), 0.01)))
