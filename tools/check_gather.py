#!/usr/bin/env python
"""Developer / evidence tool (N GPUs, torchrun): the fused NMS + peer-memory gather against the NCCL all-gather path --
identical results, and the device time of both.
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_gather.py"""
import datetime
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidardetection_b200 import sharded, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
b, s = synth.cfg2(64, 4096, seed=202 + 1000 * rank)
tb, ts = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(fused, local_inputs=True, boxes=tb, scores=ts):
    return sharded.nms_batched_sharded(boxes, scores, 0.01, max_keep=500, local_inputs=local_inputs, fused=fused)


k0, n0 = run(False)
k1, n1 = run(True)
ok = torch.equal(k0, k1.clone()) and torch.equal(n0, n1.clone())
# strong form: replicated inputs, 64 problems in total
gb, gs = synth.cfg2(64, 4096, seed=202)
gtb, gts = torch.from_numpy(gb).cuda(), torch.from_numpy(gs).cuda()
ks0, ns0 = run(False, False, gtb, gts)
ks1, ns1 = run(True, False, gtb, gts)
ok = ok and torch.equal(ks0, ks1.clone()) and torch.equal(ns0, ns1.clone())
for it in range(5):  # buffer reuse
    k2, n2 = run(True)
    ok = ok and torch.equal(k0, k2) and torch.equal(n0, n2)


def local_only():
    return sharded.nms_batched_sharded(tb, ts, 0.01, max_keep=500, local_inputs=True, gather=False)


def timed(fused):
    if fused is None:  # this rank's frames only, nothing gathered: what the gather costs is the difference
        global run
        run_saved, run = run, (lambda f: local_only())
        try:
            return timed(True)
        finally:
            run = run_saved
    for _ in range(5):
        run(fused)
    dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(30):
        flush.zero_()
        flush.zero_()
        flush.zero_()  # (keeps the GPU busy while Python queues the step: the events time the device, not the launch latency)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run(fused)
        e.record()
        e.synchronize()
        ms += s.elapsed_time(e)
    t = torch.tensor([ms / 30], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


t_nccl, t_fused, t_local = timed(False), timed(True), timed(None)
okt = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: fused == nccl results: {bool(okt[0])};  ms/step (max over ranks, 64 frames x 4096 per rank): nccl all-gather {t_nccl:.4f}, fused peer-memory gather {t_fused:.4f}, no gather {t_local:.4f}")
dist.destroy_process_group()
sys.exit(0 if bool(okt[0]) else 1)
