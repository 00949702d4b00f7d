"""Drop-in for pcdet/ops/iou3d_nms/iou3d_nms_utils.py (reference lines cited per function).

Same names, signatures, return types and devices; the computation goes through the C ABI of
liblidargeom.so (include/lidargeom.h) on the caller's current CUDA stream.  The *_batched functions
are additions for callers that want to drop the reference's per-frame / per-class Python loops.
"""
import torch

from ... import _lib
from ...utils import common_utils


def _cuda_f32(x, cols):
    assert x.is_cuda, "expected a CUDA tensor"
    assert x.dim() == 2 and x.shape[1] == cols
    return x.contiguous().float() if x.dtype != torch.float32 else x.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _iou_call(fn_name, boxes_a, boxes_b, flags=_lib.LG_FLAG_NONE, out=None):
    a, b = _cuda_f32(boxes_a, 7), _cuda_f32(boxes_b, 7)
    n, m = a.shape[0], b.shape[0]
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=a.device)
    if n == 0 or m == 0:
        return out
    L = _lib.lib()
    with torch.cuda.device(a.device):
        ws_bytes = L.lg_iou_workspace_bytes(n, m)
        ws = _workspace(ws_bytes, a.device)
        rc = getattr(L, fn_name)(_lib.ptr(a), n, _lib.ptr(b), m, _lib.ptr(out), out.stride(0), _lib.ptr(ws), ws.numel(),
                                 flags, _lib.stream_ptr(a.device))
    _lib.check(rc, fn_name)
    return out


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:12-28.  CPU tensors / numpy in, same kind out.

    The reference runs a single-threaded double loop on the host (iou3d_cpu.cpp:232-252).  Here the
    matrix is computed on the B200 with the un-contracted arithmetic of that CPU build
    (LG_FLAG_STRICT_FP32) and copied back; there is no host fallback.
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]
    Returns:
        ans_iou: (N, M)
    """
    boxes_a, is_numpy = common_utils.check_numpy_to_torch(boxes_a)
    boxes_b, is_numpy = common_utils.check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    _lib.require_usable_cuda('boxes_bev_iou_cpu')
    dev = torch.device('cuda', torch.cuda.current_device())
    ans = _iou_call('lg_boxes_iou_bev', boxes_a.to(dev), boxes_b.to(dev), flags=_lib.LG_FLAG_STRICT_FP32)
    ans_iou = ans.cpu().to(boxes_a.dtype)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms_utils.py:31-45.
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]
    Returns:
        ans_iou: (N, M)
    """
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _iou_call('lg_boxes_iou_bev', boxes_a, boxes_b)


def boxes_overlap_bev(boxes_a, boxes_b):
    """The extension entry boxes_overlap_bev_gpu (iou3d_nms.cpp:49-68) as a function: (N, M) overlap areas."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _iou_call('lg_boxes_overlap_bev', boxes_a, boxes_b)


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:48-81, fused into one pass (height overlap, volumes and the quotient are
    evaluated in the pair kernel with the same individually-rounded operations as the torch ops).
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]
    Returns:
        ans_iou: (N, M)
    """
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _iou_call('lg_boxes_iou3d', boxes_a, boxes_b)


_REDUCE_KINDS = {'overlap_bev': 0, 'iou_bev': 1, 'iou3d': 2}


def boxes_iou_max(boxes_a, boxes_b, kind='iou3d', rows=True, cols=False, flags=_lib.LG_FLAG_NONE):
    """Row / column (max, argmax) of the N x M matrix `kind` without materialising it.

    What the callers of boxes_iou3d_gpu go on to compute -- `torch.max(iou3d, dim=1)` (proposal_target_layer.py:107),
    argmax over both axes (axis_aligned_target_assigner.py:150-169) -- in the same pass as the IoU; at 200k x 200k the
    matrix would be 160 GB.  Values are bit-identical to the matrix entry points', argmax is the lowest index of the
    maximum (torch.max's convention).
    Returns (row_max (N,), row_argmax (N,)) and / or (col_max (M,), col_argmax (M,)), in that order.
    """
    a, b = _cuda_f32(boxes_a, 7), _cuda_f32(boxes_b, 7)
    n, m = a.shape[0], b.shape[0]
    dev = a.device
    rmax = torch.zeros(n, dtype=torch.float32, device=dev) if rows else None
    rarg = torch.zeros(n, dtype=torch.int64, device=dev) if rows else None
    cmax = torch.zeros(m, dtype=torch.float32, device=dev) if cols else None
    carg = torch.zeros(m, dtype=torch.int64, device=dev) if cols else None
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _workspace(L.lg_iou_reduce_workspace_bytes(n, m), dev)
        rc = L.lg_boxes_iou_reduce(_lib.ptr(a), n, _lib.ptr(b), m, _REDUCE_KINDS[kind], _lib.ptr(rmax), _lib.ptr(rarg), _lib.ptr(cmax),
                                   _lib.ptr(carg), _lib.ptr(ws), ws.numel(), flags, _lib.stream_ptr(dev))
    _lib.check(rc, 'lg_boxes_iou_reduce')
    out = ()
    if rows:
        out += (rmax, rarg)
    if cols:
        out += (cmax, carg)
    return out


def _nms_buffers(fn_name, P, N, dev, flags=_lib.LG_FLAG_NONE, max_keep=None):
    """outputs + scratch of one batched NMS call (allocated before the first launch of a step so that the launches follow
    each other without host work in between)"""
    keep = torch.empty((P, N if max_keep is None else min(int(max_keep), N)), dtype=torch.int64, device=dev)
    num = torch.empty((P,), dtype=torch.int32, device=dev)  # written for every problem by the kernels (no fill launch in the step)
    ws = None
    if P > 0 and N > 0:
        with torch.cuda.device(dev):
            ws = _workspace(_lib.lib().lg_nms_workspace_bytes_ex(P, N, 1 if 'normal' in fn_name else 0, flags), dev)
    return keep, num, ws


def _nms_call(fn_name, boxes, order, counts, thresh, flags=_lib.LG_FLAG_NONE, buffers=None, max_keep=None):
    """boxes (P, N, 7) cuda f32 contiguous; order (P, N) int64 or None; counts (P,) int32 or None.
    max_keep: NMS_POST_MAXSIZE -- keep comes back as (P, min(max_keep, N)) and the kernels stop once a row is full."""
    P, N = boxes.shape[0], boxes.shape[1]
    dev = boxes.device
    keep, num, ws = buffers if buffers is not None else _nms_buffers(fn_name, P, N, dev, flags, max_keep)
    if P == 0 or N == 0 or keep.shape[1] == 0:
        return keep, num.zero_()
    L = _lib.lib()
    mk = keep.shape[1] if max_keep is None else min(int(max_keep), keep.shape[1])  # a wider keep keeps its pitch; columns >= mk stay unwritten
    with torch.cuda.device(dev):
        rc = L.lg_nms_batched_ex(_lib.ptr(boxes), _lib.ptr(order), _lib.ptr(counts), P, N, float(thresh), 1 if 'normal' in fn_name else 0,
                                 mk, keep.stride(0), _lib.ptr(ws), ws.numel(), _lib.ptr(keep), _lib.ptr(num), flags,
                                 _lib.stream_ptr(dev))
    _lib.check(rc, fn_name)
    return keep, num


LG_SELECT_MAX_K = 4096


_SM_COUNT = {}


def _sm_count(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _argsort_desc(scores):
    """(P, N) scores -> (P, N) int64 indices in descending score (the wrapper's `scores.sort(descending=True)[1]`,
    iou3d_nms_utils.py:92).  Up to 4096 float32 scores per problem: lg_select_topk (one CTA per problem, stable radix sort in shared
    memory) when the batch has at most two problems per SM; otherwise torch's stable segmented sort.  ONE ordering rule on both
    branches: equal scores by ascending index, -0.0 == +0.0, NaN first -- so nms_gpu(frame) and nms_gpu_batched(frames) agree
    whatever the batch size."""
    P, N = scores.shape
    if not (scores.is_cuda and scores.dtype == torch.float32 and 0 < N <= LG_SELECT_MAX_K and 0 < P <= 2 * _sm_count(scores.device)):
        # one CTA per problem: a batch of thousands of small problems is better served by torch's segmented sort (measured on nms_cfg5)
        return scores.sort(dim=1, descending=True, stable=True)[1].contiguous()
    L = _lib.lib()
    dev = scores.device
    sc = scores.contiguous()
    order = torch.empty((P, N), dtype=torch.int64, device=dev)
    cnt = torch.empty((P,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(L.lg_select_workspace_bytes(P, N), dev)
        rc = L.lg_select_topk(_lib.ptr(sc), P, N, N, 0.0, 0, None, 0, 0, 1, _lib.ptr(order), _lib.ptr(cnt), None, _lib.ptr(ws), ws.numel(), 0,
                              _lib.stream_ptr(dev))
    _lib.check(rc, 'lg_select_topk')
    return order


def _nms_single(fn_name, boxes, scores, thresh, pre_maxsize):
    assert boxes.shape[1] == 7
    order = _argsort_desc(scores.reshape(1, -1))[0]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    if pre_maxsize is not None and order.shape[0] < boxes.shape[0]:
        # reference path: gather first, map back afterwards (iou3d_nms_utils.py:93-99)
        b = _cuda_f32(boxes[order], 7).unsqueeze(0)
        keep, num = _nms_call(fn_name, b, None, None, thresh)
        return order[keep[0, :int(num.item())]].contiguous(), None
    b = _cuda_f32(boxes, 7).unsqueeze(0)
    keep, num = _nms_call(fn_name, b, order.contiguous().unsqueeze(0), None, thresh)
    return keep[0, :int(num.item())].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """iou3d_nms_utils.py:84-99.
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (LongTensor[cuda] of kept indices into `boxes`, descending score; None)
    """
    return _nms_single('lg_nms_rotated_batched', boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """iou3d_nms_utils.py:102-116 (axis-aligned BEV IoU, heading ignored; no pre_maxsize).
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (LongTensor[cuda] of kept indices into `boxes`, descending score; None)
    """
    return _nms_single('lg_nms_normal_batched', boxes, scores, thresh, None)


def _nms_batched(fn_name, boxes, scores, thresh, counts, flags=_lib.LG_FLAG_NONE, max_keep=None, keep_out=None):
    assert boxes.dim() == 3 and boxes.shape[2] == 7 and scores.shape == boxes.shape[:2]
    b = boxes.contiguous().float()
    if counts is not None:
        # invalid (padding) rows sort last
        idx = torch.arange(b.shape[1], device=b.device).unsqueeze(0)
        scores = scores.masked_fill(idx >= counts.to(b.device).unsqueeze(1), float('-inf'))
        counts = counts.to(device=b.device, dtype=torch.int32).contiguous()
    keep, num, ws = _nms_buffers(fn_name, b.shape[0], b.shape[1], b.device, flags, max_keep)
    if keep_out is not None:  # caller-owned keep (any row pitch, e.g. a slice of a send buffer): written in place
        assert keep_out.dtype == torch.int64 and keep_out.shape == keep.shape and keep_out.stride(1) == 1 and keep_out.device == b.device
        keep = keep_out
    order = _argsort_desc(scores)
    return _nms_call(fn_name, b, order, counts, thresh, flags, (keep, num, ws), max_keep)


def nms_gpu_batched(boxes, scores, thresh, counts=None, full_mask=False, max_keep=None, keep_out=None):
    """P independent rotated-NMS problems in two launches (score sort, lazy NMS incl. its records) and no host sync.
    :param boxes: (P, N, 7), :param scores: (P, N), :param counts: optional (P,) valid boxes per problem
    :param full_mask: materialise the reference's N x N/64 suppression mask and sweep it (three launches)
        instead of evaluating kept rows only; the keep lists are identical
    :param max_keep: NMS_POST_MAXSIZE (model_nms_utils.py:20): only the first max_keep kept boxes are wanted; keep is then
        (P, min(max_keep, N)) and the kernel stops once a problem's row is full
    :param keep_out: optional caller-owned int64 (P, K) tensor (unit column stride, any row pitch) to receive keep
    :return: keep (P, N) int64 indices into each problem's boxes, padded with -1; num_keep (P,) int32
    """
    flags = _lib.LG_FLAG_NMS_FULL_MASK if full_mask else _lib.LG_FLAG_NONE
    return _nms_batched('lg_nms_rotated_batched', boxes, scores, thresh, counts, flags, max_keep, keep_out)


def nms_normal_gpu_batched(boxes, scores, thresh, counts=None, max_keep=None, keep_out=None):
    """Batched axis-aligned NMS; see nms_gpu_batched."""
    return _nms_batched('lg_nms_normal_batched', boxes, scores, thresh, counts, max_keep=max_keep, keep_out=keep_out)


def nms_gpu_gather(boxes, scores, thresh, max_keep, peer_ptrs, row0, counts=None):
    """Rotated NMS of this rank's P problems with the results written, by the NMS kernel itself, into row row0 + p of a packed
    (rows, 1 + max_keep) int64 buffer on EVERY rank (lg_nms_rotated_gather): column 0 the count, then the kept indices, -1 padded.
    peer_ptrs: sequence of the buffers' addresses in this process (torch symmetric memory `buffer_ptrs`), own rank included.
    The caller synchronises the ranks afterwards (lidardetection_b200.sharded does: one symmetric-memory barrier)."""
    import ctypes as C

    assert boxes.dim() == 3 and boxes.shape[2] == 7 and scores.shape == boxes.shape[:2] and boxes.is_cuda
    b = boxes.contiguous().float()
    P, N = b.shape[0], b.shape[1]
    if P == 0:
        return
    if counts is not None:
        idx = torch.arange(N, device=b.device).unsqueeze(0)
        scores = scores.masked_fill(idx >= counts.to(b.device).unsqueeze(1), float('-inf'))
        counts = counts.to(device=b.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    with torch.cuda.device(b.device):
        ws = _workspace(L.lg_nms_workspace_bytes_ex(P, N, 0, 0), b.device)
    order = _argsort_desc(scores)
    arr = (C.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    with torch.cuda.device(b.device):
        rc = L.lg_nms_rotated_gather(_lib.ptr(b), _lib.ptr(order), _lib.ptr(counts), P, N, float(thresh), int(max_keep), _lib.ptr(ws), ws.numel(),
                                     arr, len(peer_ptrs), int(row0), None, _lib.LG_FLAG_NONE, _lib.stream_ptr(b.device))
    _lib.check(rc, 'lg_nms_rotated_gather')


_SIDE_STREAMS = {}


def nms_gpu_batched_from_host(h_boxes, h_scores, thresh, counts=None, max_keep=None, device=None):
    """nms_gpu_batched for inputs that still live in (pinned) host memory -- the serving case: the scores (1/7 of the bytes) are
    uploaded first and sorted while the boxes are still crossing PCIe on a second stream; the NMS kernel then waits for both.
    Same results as nms_gpu_batched(h_boxes.cuda(), h_scores.cuda(), ...).
    :return: keep, num_keep on the device (current stream)"""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    cur = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
    assert h_boxes.dim() == 3 and h_boxes.shape[2] == 7 and h_scores.shape == h_boxes.shape[:2]
    sc = h_scores.to(dev, non_blocking=True)  # current stream: the scores go first ...
    side.wait_stream(cur)                     # (the side stream only has to respect what was queued before this call)
    with torch.cuda.stream(side):
        b = h_boxes.to(dev, dtype=torch.float32, non_blocking=True)  # ... the boxes follow on the side stream
    sc = sc.float()
    if counts is not None:
        idx = torch.arange(sc.shape[1], device=dev).unsqueeze(0)
        sc = sc.masked_fill(idx >= counts.to(dev).unsqueeze(1), float('-inf'))
        counts = counts.to(device=dev, dtype=torch.int32).contiguous()
    fn = 'lg_nms_rotated_batched'
    buffers = _nms_buffers(fn, b.shape[0], b.shape[1], dev, _lib.LG_FLAG_NONE, max_keep)
    order = _argsort_desc(sc)                 # overlaps the box upload
    cur.wait_stream(side)
    b.record_stream(cur)
    return _nms_call(fn, b.contiguous(), order, counts, thresh, _lib.LG_FLAG_NONE, buffers, max_keep)


class HostNmsPipeline:
    """Serving loop for batches of frames that arrive in (pinned) host memory, one batch after the other
    (post-processing of a detector that runs elsewhere, or offline evaluation from saved predictions).

    submit() queues  upload -> score sort -> rotated NMS -> download  of one batch on the pipeline's own three streams and returns at
    once; result(ticket) waits for that batch only.  `depth` batches are in flight, each with its own device and pinned result
    buffers: the boxes of batch k + 1 cross PCIe while batch k is in the kernels (on nms_cfg2 the upload, 8.4 MB, takes about as
    long as the kernels do).  Every batch is copied host -> device and its result device -> host; nothing is cached between
    batches.  Results are those of nms_gpu_batched(h_boxes.cuda(), h_scores.cuda(), thresh, max_keep=max_keep).

        pipe = HostNmsPipeline(64, 4096, 0.01, max_keep=500)
        t = pipe.submit(h_boxes, h_scores)       # returns immediately
        ...
        h_keep, h_num = pipe.result(t)           # pinned (P, max_keep) int64 / (P,) int32; valid until `depth` more submits
    """

    def __init__(self, num_problems, num_boxes, thresh, max_keep=None, depth=2, device=None):
        assert depth >= 1 and num_problems > 0 and num_boxes > 0
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.P, self.N, self.thresh, self.depth = int(num_problems), int(num_boxes), float(thresh), int(depth)
        self.K = self.N if max_keep is None else min(int(max_keep), self.N)
        self.up, self.comp, self.down = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.slots = []
        for _ in range(self.depth):
            self.slots.append({
                'boxes': torch.empty((self.P, self.N, 7), dtype=torch.float32, device=self.dev),
                'scores': torch.empty((self.P, self.N), dtype=torch.float32, device=self.dev),
                'keep': torch.empty((self.P, self.K), dtype=torch.int64, device=self.dev),
                'h_keep': torch.empty((self.P, self.K), dtype=torch.int64).pin_memory(),
                'h_num': torch.empty((self.P,), dtype=torch.int32).pin_memory(),
                'ev_up': torch.cuda.Event(), 'ev_nms': torch.cuda.Event(), 'ev_done': torch.cuda.Event(), 'ticket': -1, 'pending': False,
            })
        self.next_ticket = 0

    def submit(self, h_boxes, h_scores):
        assert tuple(h_boxes.shape) == (self.P, self.N, 7) and tuple(h_scores.shape) == (self.P, self.N)
        assert h_boxes.dtype == torch.float32 and h_scores.dtype == torch.float32 and not h_boxes.is_cuda
        t = self.next_ticket
        s = self.slots[t % self.depth]
        if s['pending']:
            raise RuntimeError('HostNmsPipeline: result() of ticket %d has not been collected (depth %d)' % (s['ticket'], self.depth))
        with torch.cuda.stream(self.up):
            # the slot's previous batch has left the kernels (its result() synchronised on ev_done): the buffers are free
            s['scores'].copy_(h_scores, non_blocking=True)
            s['boxes'].copy_(h_boxes, non_blocking=True)
            s['ev_up'].record(self.up)
        with torch.cuda.stream(self.comp):
            self.comp.wait_event(s['ev_up'])
            keep, num = nms_gpu_batched(s['boxes'], s['scores'], self.thresh, max_keep=self.K, keep_out=s['keep'])
            s['ev_nms'].record(self.comp)
        with torch.cuda.stream(self.down):  # the download does not hold up the next batch's kernels
            self.down.wait_event(s['ev_nms'])
            s['h_keep'].copy_(keep, non_blocking=True)
            s['h_num'].copy_(num, non_blocking=True)
            num.record_stream(self.down)  # allocated on the compute stream, last read here
            s['ev_done'].record(self.down)
        s['ticket'], s['pending'] = t, True
        self.next_ticket += 1
        return t

    def result(self, ticket):
        s = self.slots[ticket % self.depth]
        if s['ticket'] != ticket or not s['pending']:
            raise RuntimeError('HostNmsPipeline: ticket %d is not in flight' % ticket)
        s['ev_done'].synchronize()
        s['pending'] = False
        return s['h_keep'], s['h_num']
