"""One-process-per-GPU drivers for the parts of the hot path that shard naturally.

    IoU matrices        rows of boxes_a split into contiguous blocks, boxes_b replicated; no data-path
                        collective.  The (N, M) result is all-gathered only on request -- at 200k x 200k
                        it is 160 GB and stays row-sharded.
    batched NMS         problems (frames x classes) dealt in contiguous blocks; keep lists gathered.
    points in boxes     frames dealt in contiguous blocks; (B, M) indices gathered.

The reference has no multi-GPU code on this path (its only distributed code is DDP and a pickle-file
merge, pcdet/utils/common_utils.py:146-227).  Collectives go through torch.distributed (NCCL over
NVLink on the B200 box, gloo in the CPU unit tests); the `compute` callables default to the CUDA ops
and are injectable so that the partition / gather logic is testable without a GPU.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous block [start, stop) of n items for `rank`: blocks of ceil(n / world) (last ones short/empty)."""
    per = (n + world_size - 1) // world_size if world_size > 0 else n
    start = min(rank * per, n)
    return start, min(start + per, n)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _gather_rows(local, n_total, group):
    """all-gather row blocks produced with shard_range (equal-size padded blocks) -> (n_total, ...)."""
    rank, world = _world(group)
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def boxes_iou_sharded(boxes_a, boxes_b, kind="iou3d", gather=False, group=None, compute=None):
    """Row-sharded N x M IoU.  Every rank passes the SAME boxes_a / boxes_b (replicated inputs, KBs-MBs).

    Returns (block, (start, stop)) with block = rows [start, stop) of the matrix computed by this rank,
    or, with gather=True, the full (N, M) matrix on every rank.
    kind: 'iou3d' | 'iou_bev' | 'overlap_bev'.
    """
    if compute is None:
        from .ops.iou3d_nms import iou3d_nms_utils as U

        compute = {"iou3d": U.boxes_iou3d_gpu, "iou_bev": U.boxes_iou_bev, "overlap_bev": U.boxes_overlap_bev}[kind]
    rank, world = _world(group)
    n = boxes_a.shape[0]
    start, stop = shard_range(n, rank, world)
    block = compute(boxes_a[start:stop], boxes_b)
    if gather:
        return _gather_rows(block, n, group)
    return block, (start, stop)


class _SymmRows:
    """Two alternating symmetric-memory (rows, cols) int64 result buffers per (group, shape, device) for the fused NMS gather:
    allocated and exchanged once (torch.distributed._symmetric_memory: CUDA peer mappings over NVLink), then reused.  A buffer
    handed out by get() is rewritten by the call after next; the barrier of the call in between orders that write after
    every rank's reads of it, provided those reads were enqueued on the stream the sharded call runs on."""
    _cache = {}

    @classmethod
    def get(cls, group, rows, cols, device):
        import torch.distributed._symmetric_memory as symm_mem

        g = group if group is not None else dist.group.WORLD
        key = (g.group_name, rows, cols, device.index)
        if key not in cls._cache:
            bufs, hdls = [], []
            for _ in range(2):
                t = symm_mem.empty((rows, cols), dtype=torch.int64, device=device)
                hdls.append(symm_mem.rendezvous(t, g))
                bufs.append(t)
            cls._cache[key] = [bufs, hdls, 0]
        e = cls._cache[key]
        i = e[2]
        e[2] ^= 1
        return e[0][i], e[1][i]


def _fused_gather_ok(boxes, world, normal, compute, blocks_equal):
    """the fused path needs: NCCL world on CUDA tensors (one NVLink domain of <= 8 ranks), the CUDA op, rotated NMS, equal blocks"""
    if compute is not None or normal or not boxes.is_cuda or not blocks_equal or world > 8:
        return False
    try:
        return dist.get_backend() == "nccl"
    except Exception:  # noqa: BLE001
        return False


def nms_batched_sharded(boxes, scores, thresh, counts=None, normal=False, gather=True, group=None, compute=None, max_keep=None,
                        local_inputs=False, fused=None):
    """Frame-sharded batched NMS.  boxes (P, N, 7), scores (P, N) replicated on every rank (local_inputs=False: every rank
    works on its contiguous block of the P problems), or -- local_inputs=True, the data-parallel inference case -- already
    this rank's own block of a global batch of world x P problems.

    max_keep: NMS_POST_MAXSIZE (model_nms_utils.py:20) -- only the first max_keep kept boxes of a problem are produced and
    gathered (the reference merges the truncated per-frame results, common_utils.py:206-227, through pickle files and two
    barriers; here it is ONE all_gather_into_tensor of a packed (block, 1 + max_keep) int64 tensor: column 0 the count, the
    rest the kept indices, which the NMS kernel writes straight into the send buffer).
    fused (default: on where it applies -- NCCL job, rotated NMS, equal blocks, <= 8 ranks of one NVLink domain): the NMS kernel
    writes its results into every rank's buffer itself and a symmetric-memory barrier replaces the collective (see below);
    keep / num_keep are then views into a reused buffer, valid until the call after next.  fused=False: the NCCL all-gather.
    Returns keep (P_total, K) int64 (-1 padded) and num_keep (P_total,) int32 for all problems (gather=True), or this
    rank's block plus its (start, stop).
    """
    user_compute = compute
    if compute is None:
        from .ops.iou3d_nms import iou3d_nms_utils as U

        compute = U.nms_normal_gpu_batched if normal else U.nms_gpu_batched
    rank, world = _world(group)
    if local_inputs:
        P = boxes.shape[0] * world
        start, stop = rank * boxes.shape[0], (rank + 1) * boxes.shape[0]
        b, sc, c = boxes, scores, counts
    else:
        P = boxes.shape[0]
        start, stop = shard_range(P, rank, world)
        b, sc, c = boxes[start:stop], scores[start:stop], (None if counts is None else counts[start:stop])
    kw = {} if max_keep is None else {"max_keep": max_keep}
    if fused is None:
        fused = True
    if gather and world > 1 and fused and not thresh < 0 and _fused_gather_ok(boxes, world, normal, user_compute, local_inputs or P % world == 0):
        # ONE kernel does the NMS and the gather: its epilogue stores every problem's packed (count, kept indices) row into the
        # result buffer of every rank through the NVLink peer mappings; a symmetric-memory barrier (signal pads, no NCCL launch)
        # then tells every rank that all rows have landed.  The result is a view into the (reused) symmetric buffer.
        from .ops.iou3d_nms import iou3d_nms_utils as U

        per = stop - start
        K = b.shape[1] if max_keep is None else min(int(max_keep), b.shape[1])
        buf, hdl = _SymmRows.get(group, world * per, 1 + K, b.device)
        U.nms_gpu_gather(b, sc, thresh, K, hdl.buffer_ptrs, rank * per, c)
        hdl.barrier(channel=0)
        return buf[:, 1:], buf[:, 0].to(torch.int32)
    if not gather or world == 1:
        keep, num = compute(b, sc, thresh, c, **kw)
        if gather:
            return keep, num
        return (keep, num), (start, stop)
    per = boxes.shape[0] if local_inputs else (P + world - 1) // world
    K = b.shape[1] if max_keep is None else min(int(max_keep), b.shape[1])
    send = torch.empty((per, 1 + K), dtype=torch.int64, device=b.device)
    if stop - start < per:  # a short (or empty) last block: pad rows, count 0
        send[stop - start:, 0] = 0
        send[stop - start:, 1:] = -1
    if stop > start:
        try:  # the CUDA op takes its keep buffer from the caller: the kernel writes into the send buffer
            keep, num = compute(b, sc, thresh, c, keep_out=send[: stop - start, 1:], **kw)
        except TypeError:  # an injected compute (CPU tests) without keep_out
            keep, num = compute(b, sc, thresh, c, **kw)
            send[: stop - start, 1:] = keep[:, :K]
        send[: stop - start, 0] = num
    recv = torch.empty((world * per, 1 + K), dtype=torch.int64, device=b.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:P, 1:], recv[:P, 0].to(torch.int32)


def points_in_boxes_sharded(points, boxes, gather=True, group=None, compute=None):
    """Frame-sharded points_in_boxes_gpu.  points (B, M, 3), boxes (B, T, 7) replicated on every rank."""
    if compute is None:
        from .ops.roiaware_pool3d import roiaware_pool3d_utils as PU

        compute = PU.points_in_boxes_gpu
    rank, world = _world(group)
    B = points.shape[0]
    start, stop = shard_range(B, rank, world)
    idx = compute(points[start:stop], boxes[start:stop])
    if gather:
        return _gather_rows(idx, B, group)
    return idx, (start, stop)


def boxes_iou_max_sharded(boxes_a, boxes_b, kind="iou3d", group=None, compute=None):
    """Row-sharded (max, argmax) of the N x M IoU matrix over both axes, the matrix itself never materialised
    (SURVEY.md 8f-4).  Every rank passes the SAME boxes_a / boxes_b.

    Rows split as in boxes_iou_sharded, so a rank's row maxima are final; the column maxima of the row shards are
    combined with ONE all-reduce(MAX) of M packed 64-bit keys (value bits << 32 | ~global row index: a larger
    value wins, equal values keep the lower row) -- the only real exchange step on this path.
    Returns (row_max, row_argmax, (start, stop)) for this rank's rows and (col_max (M,), col_argmax (M,)) global.
    """
    if compute is None:
        from .ops.iou3d_nms import iou3d_nms_utils as U

        def compute(a, b):
            return U.boxes_iou_max(a, b, kind=kind, rows=True, cols=True)
    rank, world = _world(group)
    n = boxes_a.shape[0]
    start, stop = shard_range(n, rank, world)
    rmax, rarg, cmax, carg = compute(boxes_a[start:stop], boxes_b)
    if world > 1:
        # IoU values are >= 0, so their float bits order like integers and the key fits a signed int64
        key = (cmax.contiguous().view(torch.int32).to(torch.int64) << 32) | (0xFFFFFFFF - (carg + start))
        if stop == start:  # a rank without rows contributes the neutral key (value 0.0, highest index)
            key = torch.zeros_like(key)
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
        cmax = (key >> 32).to(torch.int32).view(torch.float32)
        carg = 0xFFFFFFFF - (key & 0xFFFFFFFF)
        carg = torch.where(cmax > 0, carg, torch.zeros_like(carg))  # an all-zero column: index 0, as torch.max of the full matrix
    return (rmax, rarg, (start, stop)), (cmax, carg)
