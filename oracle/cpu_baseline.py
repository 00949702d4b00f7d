"""Timed CPU baseline for bench.py (the only place outside tests/ and smoke() that executes oracle/).

kind = "reference": the reference's own compiled CPU functions (oracle/_ref: boxes_iou_bev_cpu,
                    points_in_boxes_cpu) in a pool of forked worker processes (they hold the GIL);
                    the pool must be created BEFORE the parent initialises CUDA;
kind = "port":      the C restatement (oracle/liblg_oracle.so, CPU flavor) in a thread pool (ctypes
                    releases the GIL).
The reference has no CPU NMS: the CPU NMS baseline is boxes_iou_bev_cpu on row blocks of the
score-sorted boxes (upper block-triangle only) followed by the host sweep of iou3d_nms.cpp:116-132.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_STATE = {}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _worker_init(kind):
    import torch

    torch.set_num_threads(1)
    if kind == "reference":
        from oracle import ref_loader

        _STATE["ref"] = ref_loader.iou3d_nms_cuda()
        _STATE["roi"] = ref_loader.roiaware_pool3d_cuda()
    from oracle import lg_oracle

    _STATE["orc"] = lg_oracle
    lg_oracle.lib()
    _STATE["kind"] = kind


def _iou_block(args):
    a, b, thresh = args
    if _STATE["kind"] == "reference":
        import torch

        out = torch.zeros(a.shape[0], b.shape[0])
        _STATE["ref"].boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), out)
        iou = out.numpy()
    else:
        iou = _STATE["orc"].boxes_iou_bev(a, b, 0)
    if thresh is None:
        return float(iou.sum())
    return np.packbits(iou > thresh, axis=1, bitorder="little")


def _pib_block(args):
    pts, boxes = args
    if _STATE["kind"] == "reference":
        import torch

        out = torch.zeros(boxes.shape[0], pts.shape[0], dtype=torch.int32)
        _STATE["roi"].points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(pts), out)
        return int(out.sum())
    return int(_STATE["orc"].points_in_boxes_mask(pts, boxes, 1e-2, 0).sum())


def _roiaware_frame(args):
    pts, rois, feat_part, feat_rpn, out, max_pts = args
    o = _STATE["orc"]
    a = o.roiaware_pool3d_forward(rois, pts, feat_part, out, max_pts, "avg")[0]
    b = o.roiaware_pool3d_forward(rois, pts, feat_rpn, out, max_pts, "max")[0]
    return float(a.sum()) + float(b.sum())


def _kitti_part(args):
    g, d = args  # float64 camera boxes of one evaluation part: BEV overlap + 3-D overlap, as eval.py:370-393 computes them
    o = _STATE["orc"]
    a = o.bev_box_overlap(g[:, [0, 2, 3, 5, 6]], d[:, [0, 2, 3, 5, 6]], -1, 0)
    b = o.d3_box_overlap(g, d, -1, 0)
    return float(a.sum()) + float(b.sum())


def _roipoint_frame(args):
    pts, boxes, feat, s = args
    return float(_STATE["orc"].roipoint_pool3d_forward(pts[None], feat[None], boxes[None], s)[0].sum())


class CpuPool:
    def __init__(self, prefer_reference=True, workers=None):
        from oracle import ref_loader

        self.kind = "reference" if (prefer_reference and ref_loader.available()) else "port"
        self.workers = workers or host_cores()
        if self.kind == "reference":
            import multiprocessing as mp

            import torch

            assert not torch.cuda.is_initialized(), "create the CPU pool before touching CUDA (fork)"
            self._pool = mp.get_context("fork").Pool(self.workers, initializer=_worker_init, initargs=(self.kind,))
            self._map = self._pool.map
        else:
            _worker_init("port")
            self._pool = ThreadPoolExecutor(self.workers)
            self._map = lambda fn, it: list(self._pool.map(fn, it))
        self._map(_warm, range(self.workers))

    def close(self):
        if self.kind == "reference":
            self._pool.close()
            self._pool.join()
        else:
            self._pool.shutdown()

    # ---- workloads -------------------------------------------------------------------------
    def nms_frame(self, boxes, scores, thresh):
        """one NMS problem on the CPU; returns (keep indices, seconds)"""
        t0 = time.perf_counter()
        order = np.argsort(-scores, kind="stable")
        b = np.ascontiguousarray(boxes[order])
        n = b.shape[0]
        blk = max(64, (n // (self.workers * 4) + 63) // 64 * 64)
        jobs = [(b[r:r + blk], b[r:], thresh) for r in range(0, n, blk)]
        parts = self._map(_iou_block, jobs)
        dead = np.zeros(n, dtype=bool)
        keep = []
        for j, r0 in enumerate(range(0, n, blk)):
            bits = np.unpackbits(parts[j], axis=1, bitorder="little")[:, : n - r0].astype(bool)
            for i in range(bits.shape[0]):
                g = r0 + i
                if not dead[g]:
                    keep.append(g)
                    row = bits[i]
                    row[: i + 1] = False  # strictly later boxes only (diagonal tile starts at t+1)
                    dead[r0:] |= row
        return order[np.asarray(keep, dtype=np.int64)], time.perf_counter() - t0

    def iou_matrix(self, a, b):
        """N x M BEV IoU on the CPU in row blocks; returns seconds"""
        t0 = time.perf_counter()
        blk = max(1, (a.shape[0] + self.workers * 4 - 1) // (self.workers * 4))
        self._map(_iou_block, [(a[r:r + blk], b, None) for r in range(0, a.shape[0], blk)])
        return time.perf_counter() - t0

    def points_mask(self, pts, boxes):
        """frames of points_in_boxes_cpu; pts (B,M,3), boxes (B,T,7); returns seconds"""
        t0 = time.perf_counter()
        self._map(_pib_block, [(pts[f], boxes[f]) for f in range(pts.shape[0])])
        return time.perf_counter() - t0


    def roiaware_frames(self, frames):
        """Part-A2 RoI-aware pooling (avg + max call) per frame; the reference has no CPU build of it: always the port"""
        t0 = time.perf_counter()
        self._map(_roiaware_frame, frames)
        return time.perf_counter() - t0

    def kitti_parts(self, parts):
        """KITTI-eval overlaps (bev + 3d) of evaluation parts; the reference's kernel is numba.cuda only: always the port"""
        t0 = time.perf_counter()
        self._map(_kitti_part, parts)
        return time.perf_counter() - t0

    def roipoint_frames(self, frames):
        t0 = time.perf_counter()
        self._map(_roipoint_frame, frames)
        return time.perf_counter() - t0


def _warm(_):
    return 0
