"""CPU-only, world_size 2 over gloo: the row / frame partition + gather logic of lidardetection_b200.sharded,
with the oracle standing in for the CUDA ops (the partitioning code is device-agnostic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidardetection_b200 import sharded, synth
from oracle import lg_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _iou_cpu(a, b):
    return torch.from_numpy(O.boxes_iou3d(a.numpy(), b.numpy(), O.FLAVOR_CUDA))


def _nms_cpu(boxes, scores, thresh, counts, max_keep=None):
    P, N = scores.shape
    K = N if max_keep is None else min(max_keep, N)
    keep = torch.full((P, K), -1, dtype=torch.int64)
    num = torch.zeros((P,), dtype=torch.int32)
    for p in range(P):
        k = O.nms(boxes[p].numpy(), scores[p].numpy(), thresh)[:K]
        keep[p, : len(k)] = torch.from_numpy(k)
        num[p] = len(k)
    return keep, num


def _pib_cpu(points, boxes):
    return torch.from_numpy(O.points_in_boxes_idx(points.numpy(), boxes.numpy(), O.FLAVOR_CUDA))


def _max_cpu(a, b):
    """stand-in for U.boxes_iou_max: (row max, row argmax, col max, col argmax) with lowest-index ties"""
    mat = _iou_cpu(a, b)
    if mat.shape[0] == 0:
        z = torch.zeros(mat.shape[1])
        return torch.zeros(0), torch.zeros(0, dtype=torch.int64), z, z.long()
    r, c = mat.max(1), mat.max(0)
    return r.values, (mat == r.values[:, None]).int().argmax(1), c.values, (mat == c.values[None, :]).int().argmax(0)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = synth.clustered_pairs(37, 11, 3)
        ta, tb = torch.from_numpy(a), torch.from_numpy(b)
        full = sharded.boxes_iou_sharded(ta, tb, gather=True, compute=_iou_cpu)
        block, (s, e) = sharded.boxes_iou_sharded(ta, tb, gather=False, compute=_iou_cpu)
        ok = torch.equal(full, _iou_cpu(ta, tb)) and torch.equal(block, full[s:e]) and (s, e) == sharded.shard_range(37, rank, world)
        boxes, scores = synth.nms_frames(5, 60, seed=8)
        keep, num = sharded.nms_batched_sharded(torch.from_numpy(boxes), torch.from_numpy(scores), 0.1, compute=_nms_cpu)
        k1, n1 = _nms_cpu(torch.from_numpy(boxes), torch.from_numpy(scores), 0.1, None)
        ok = ok and torch.equal(keep, k1) and torch.equal(num, n1)
        # NMS_POST_MAXSIZE: only the first 7 kept boxes are produced and gathered (one packed all-gather)
        keep7, num7 = sharded.nms_batched_sharded(torch.from_numpy(boxes), torch.from_numpy(scores), 0.1, compute=_nms_cpu, max_keep=7)
        ok = ok and torch.equal(keep7, k1[:, :7]) and torch.equal(num7, torch.clamp(n1, max=7)) and keep7.shape == (5, 7)
        # data-parallel form: every rank brings its own frames; the result covers world x local frames in rank order
        lb, ls = synth.nms_frames(3, 40, seed=20 + rank)
        keepl, numl = sharded.nms_batched_sharded(torch.from_numpy(lb), torch.from_numpy(ls), 0.1, compute=_nms_cpu, max_keep=9, local_inputs=True)
        ok = ok and keepl.shape == (3 * world, 9)
        for r in range(world):
            rb, rs = synth.nms_frames(3, 40, seed=20 + r)
            kr, nr = _nms_cpu(torch.from_numpy(rb), torch.from_numpy(rs), 0.1, None, 9)
            ok = ok and torch.equal(keepl[3 * r:3 * r + 3], kr) and torch.equal(numl[3 * r:3 * r + 3], nr)
        pts, rois = synth.cfg3(n_frames=3, n_points=200, n_rois=12, seed=4)
        idx = sharded.points_in_boxes_sharded(torch.from_numpy(pts), torch.from_numpy(rois), compute=_pib_cpu)
        ok = ok and torch.equal(idx, _pib_cpu(torch.from_numpy(pts), torch.from_numpy(rois)))
        # fused maxima: rows local, columns combined with one all-reduce(MAX) of packed keys
        a2, b2 = synth.clustered_pairs(41, 29, 5)
        b2[7] = [500, 500, 0, 1, 1, 1, 0]  # a column that overlaps nothing: (0.0, index 0)
        ta2, tb2 = torch.from_numpy(a2), torch.from_numpy(b2)
        (rmax, rarg, (s2, e2)), (cmax, carg) = sharded.boxes_iou_max_sharded(ta2, tb2, compute=_max_cpu)
        wr, wra, wc, wca = _max_cpu(ta2, tb2)
        ok = ok and torch.equal(rmax, wr[s2:e2]) and torch.equal(rarg, wra[s2:e2]) and torch.equal(cmax, wc) and torch.equal(carg, wca)
        ok = ok and float(cmax[7]) == 0.0 and int(carg[7]) == 0
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_partition_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(170)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_single_process_path_needs_no_process_group():
    a, b = synth.clustered_pairs(9, 4, 2)
    full = sharded.boxes_iou_sharded(torch.from_numpy(a), torch.from_numpy(b), gather=True, compute=_iou_cpu)
    assert np.array_equal(full.numpy(), O.boxes_iou3d(a, b, O.FLAVOR_CUDA))
