#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N B200s of one node; prints ONE JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`; a "step" is one pass of the op over one resident batch of synthetic input):
    nms_cfg2  (default) SECOND KITTI post-processing: rotated nms_gpu, 4096 boxes/frame, thresh 0.01,
              64 frames per GPU (weak scaling: frames are independent problems, no data-path collective;
              every rank keeps the keep lists of its own frames, as the reference's DDP evaluation does).  metric: frames/s.
    nms_cfg5  NuScenes CBGS multi-head: 1000 boxes x 10 classes x 256 frames, thresh 0.2.  metric: problems/s.
    iou_dense FP32-roofline microbench: 16384 x 16384 all-overlapping pairs.  metric: Gpairs/s.
    iou_cfg1  PointPillars anchors x GT, 321,408 x 20 boxes_iou_bev.  metric: Gpairs/s.
    iou_cfg4  Waymo-scale boxes_iou3d, 200k x 200k row-sharded (each rank owns 200k/8 = 25,000 rows
              = 20 GB of output whatever N is: weak scaling; strong-scaling numbers follow by division).
    pib_cfg3  PV-RCNN points_in_boxes_gpu, 16,384 points x 100 ROIs, 4096 frames per GPU.  metric: frames/s.
    post_cfg2 (next row 8f-1) the post-processing front end around nms_cfg2: 70,400 candidates -> score threshold -> top 4096
              -> NMS -> 500, 64 frames in one batched call.  metric: frames/s.
    iou_max_cfg4 (next row 8f-4) per-box best match of a 25,000-row shard against 200,000 boxes, matrix never materialised;
              column maxima all-reduced over NCCL at N > 1.  metric: Gpairs/s.

JSON keys follow the driver's contract: value = whole-job throughput with inputs resident in HBM (CUDA
events per step, summed, max over ranks); e2e = same metric through the public Python API from pinned
HOST buffers with H2D and D2H inside the timed region; roofline = dominant kernel's algorithmic work /
its own CUDA-event duration against a measured peak; cpu_baseline = the reference's CPU path on this
box's host cores over a bounded sample.  `--impl reference` times only that CPU path (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2


def l2_flush(flush):
    """Evict the L2 between timed iterations (outside the CUDA events).  Six passes over the 256 MB buffer: one evicts the cache; the
    others keep the GPU busy for ~0.25 ms, so that Python has queued the step's launches by the time the start event fires and
    the events time the GPU's work on the step, not the host's launch latency (a busy host core -- the clock sampler's thread on
    rank 0, a neighbour's process -- otherwise adds it: seen as 0.23 instead of 0.18 ms on a step of two launches, and as
    0.22 instead of 0.19 ms on the multi-GPU step, whose Python side is longer)."""
    for _ in range(6):
        flush.zero_()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def workload_config(name):
    """the `config` of a line: the workload and the cache discipline of the GPU arm; the same dict in both arms"""
    return {"workload": name, "l2": "256 MB L2 flush between timed iterations of the GPU arm (outside the events)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)", 1965.0


def ncu_traffic(workload):
    """DRAM bytes (read + write) per launch of the workload's dominant kernel, from the committed `ncu --set full` capture at
    the bench shape (profiles/rNN_ncu_traffic.json, written by tools/make_profiles.py --traffic); None when there is none."""
    rec = None
    for tag in ("r02", "r01"):  # the latest round's capture of the kernel, else the previous one's
        try:
            with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")) as f:
                rec = json.load(f).get(workload)
        except (OSError, ValueError):
            rec = None
        if rec:
            break
    if not rec:
        return None, None
    return rec["dram_bytes_read"] + rec["dram_bytes_write"], f"{rec['kernel']}: dram__bytes_read.sum + dram__bytes_write.sum, {rec['how']} ({rec['report']})"


def fp32_peak_tflops(torch):
    """unrolled-FFMA microbenchmark (tools/peak_fp32.cu): measured non-tensor FP32 peak of this GPU"""
    import ctypes as C

    path = os.path.join(ROOT, "tools", "libpeakfp32.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.peakfp32_launch.restype = C.c_double
    lib.peakfp32_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    sink = torch.zeros(148 * 16 * 256, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 0.0
    for chains in (8, 4):
        for _ in range(4):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            flops = lib.peakfp32_launch(C.c_void_p(sink.data_ptr()), 2048, 148 * 16, chains, st)
            e.record()
            e.synchronize()
            if flops > 0:
                best = max(best, flops / (s.elapsed_time(e) * 1e-3) / 1e12)
    return best


# ------------------------------------------------------------------------------------------------
class Workload:
    """inputs resident on the device + pinned host copies; step() = device hot path; e2e_step() = host->host"""
    name = metric = unit = dtype = ""
    launches_per_step = 0

    def host_inputs(self):  # numpy arrays (for the CPU baseline sample)
        raise NotImplementedError


class NmsWorkload(Workload):
    dtype = "f32"

    def __init__(self, torch, which, rank, world=1):
        from lidardetection_b200 import synth

        self.torch = torch
        self.which = which
        if which == "nms_cfg2":
            self.boxes_np, self.scores_np = synth.cfg2(64, 4096, seed=synth.SEEDS["cfg2"] + 1000 * rank)
            self.thresh, self.name = 0.01, "SECOND KITTI post-processing: rotated nms_gpu 4096 boxes/frame, thresh 0.01, 64 frames per GPU"
            self.metric, self.unit = "rotated NMS frames/s (4096 boxes/frame)", "frames/s"
            self.post = 500  # NMS_POST_MAXSIZE of the config (second.yaml:94-99, "4096 -> 500"): keep comes back as (64, 500)
        else:
            b, s = synth.cfg5(256, 10, 1000, seed=synth.SEEDS["cfg5"] + 1000 * rank)
            self.boxes_np, self.scores_np = b.reshape(-1, 1000, 7), s.reshape(-1, 1000)
            self.thresh, self.name = 0.2, "NuScenes CBGS multi-head NMS: 10 classes x 1000 boxes x 256 frames per GPU, thresh 0.2"
            self.metric, self.unit = "rotated NMS problems/s (1000 boxes/problem)", "problems/s"
            self.post = 83  # cbgs_second_multihead.yaml:196-206
        self.world = world
        self.multi_gpu_note = ("frames sharded by rank (weak scaling: every rank its own frames); the truncated keep lists + counts of ALL ranks are "
                               "gathered every step inside the timed region (lidardetection_b200.sharded.nms_batched_sharded): the NMS kernel's epilogue "
                               "stores each problem's packed (count, kept indices) row into every rank's result buffer over NVLink peer memory "
                               "(lg_nms_rotated_gather, torch symmetric memory) and one symmetric-memory barrier orders the ranks -- no NCCL launch on "
                               "the data path (fused=False selects one NCCL all_gather_into_tensor instead; NCCL carries the timing barrier and the "
                               "max-over-ranks); `strong` = the config's fixed total split over N")
        self.units = self.boxes_np.shape[0]
        self.boxes = torch.from_numpy(self.boxes_np).cuda()
        self.scores = torch.from_numpy(self.scores_np).cuda()
        self.h_boxes = torch.from_numpy(self.boxes_np).pin_memory()
        self.h_scores = torch.from_numpy(self.scores_np).pin_memory()
        # select_topk_kernel (the score sort; torch's segmented sort for batches of more than 296 problems) + nms_lazy_kernel
        self.launches_per_step = 2 if self.units <= 296 else 1
        self.h2d = self.h_boxes.numel() * 4 + self.h_scores.numel() * 4
        self.d2h = 0  # set by e2e_step from the tensors it copies

    def _nms(self, boxes, scores):
        """N = 1: the batched op.  N > 1: the same through lidardetection_b200.sharded -- every rank runs its own frames and
        the truncated keep lists + counts of ALL ranks are gathered inside the timed region (fused into the NMS kernel over
        NVLink peer memory + one symmetric-memory barrier; see sharded.nms_batched_sharded)."""
        if self.world > 1:
            from lidardetection_b200 import sharded

            return sharded.nms_batched_sharded(boxes, scores, self.thresh, max_keep=self.post, local_inputs=True)
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        return U.nms_gpu_batched(boxes, scores, self.thresh, max_keep=self.post)

    def step(self):
        return self._nms(self.boxes, self.scores)

    def local_step(self):  # the same work without the gather
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        return U.nms_gpu_batched(self.boxes, self.scores, self.thresh, max_keep=self.post)

    def e2e_step(self):
        torch = self.torch
        if self.world == 1:
            # the public host-input entry point: score upload + sort overlap the box upload (second stream)
            from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

            keep, num = U.nms_gpu_batched_from_host(self.h_boxes, self.h_scores, self.thresh, max_keep=self.post)
        else:
            b = self.h_boxes.cuda(non_blocking=True)
            s = self.h_scores.cuda(non_blocking=True)
            keep, num = self._nms(b, s)
        # the result a caller reads: counts and the (frames, NMS_POST_MAXSIZE) keep lists, into pinned buffers, ONE synchronisation
        if getattr(self, "h_keep", None) is None or self.h_keep.shape != keep.shape:
            self.h_keep = torch.empty(keep.shape, dtype=keep.dtype).pin_memory()
            self.h_num = torch.empty(num.shape, dtype=num.dtype).pin_memory()
        self.h_keep.copy_(keep, non_blocking=True)
        self.h_num.copy_(num, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h = self.h_num.numel() * 4 + self.h_keep.numel() * 8  # counted from the tensors copied
        return self.h_keep, self.h_num

    def e2e_pipelined(self, steps, depth=4):
        """the serving loop of the public API (HostNmsPipeline): every step is submitted from pinned host memory (its own H2D) and
        its result read back into pinned host memory (its own D2H); `depth` steps are in flight, so the upload of step k + 1 crosses
        PCIe while step k is in the kernels.  Returns the wall time of `steps` complete steps (first submit -> last result)."""
        import time

        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        torch = self.torch
        if self.world > 1:
            return self._e2e_pipelined_sharded(steps)
        pipe = U.HostNmsPipeline(self.units, self.boxes_np.shape[1], self.thresh, max_keep=self.post, depth=depth)
        for _ in range(3):  # warm-up: buffers, streams, allocator
            pipe.result(pipe.submit(self.h_boxes, self.h_scores))
        torch.cuda.synchronize()
        inflight = []
        t0 = time.perf_counter()
        for _ in range(steps):
            if len(inflight) == depth:
                hk, hn = pipe.result(inflight.pop(0))
            inflight.append(pipe.submit(self.h_boxes, self.h_scores))
        for t in inflight:
            hk, hn = pipe.result(t)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        self.d2h = hk.numel() * 8 + hn.numel() * 4
        return dt

    def _e2e_pipelined_sharded(self, steps):
        """N > 1: the same serving loop around sharded.nms_batched_sharded (every rank uploads its own frames, runs the NMS whose epilogue
        gathers all ranks' keep lists, downloads the gathered lists).  Two streams, two steps in flight: the upload of step k + 1
        crosses PCIe while step k is in the kernels; the download of step k stays on the kernels' stream, BEFORE step k + 1 -- the
        gathered result lives in one of two alternating symmetric buffers that the peers' kernels of step k + 2 write, and a peer
        gets there only after this rank has passed the barrier of step k + 1."""
        import time

        torch = self.torch
        dev = self.boxes.device
        up, comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        slots = [{"boxes": torch.empty_like(self.boxes), "scores": torch.empty_like(self.scores), "ev_up": torch.cuda.Event(),
                  "ev_done": torch.cuda.Event(), "h_keep": None, "h_num": None} for _ in range(2)]

        def submit(k):
            sl = slots[k % 2]
            with torch.cuda.stream(up):  # (the slot's previous step was collected by result(k - 2): its device inputs are free)
                sl["scores"].copy_(self.h_scores, non_blocking=True)
                sl["boxes"].copy_(self.h_boxes, non_blocking=True)
                sl["ev_up"].record(up)
            with torch.cuda.stream(comp):
                comp.wait_event(sl["ev_up"])
                keep, num = self._nms(sl["boxes"], sl["scores"])
                if sl["h_keep"] is None:
                    sl["h_keep"] = torch.empty(keep.shape, dtype=keep.dtype).pin_memory()
                    sl["h_num"] = torch.empty(num.shape, dtype=num.dtype).pin_memory()
                sl["h_keep"].copy_(keep, non_blocking=True)
                sl["h_num"].copy_(num, non_blocking=True)
                sl["ev_done"].record(comp)

        def result(k):
            sl = slots[k % 2]
            sl["ev_done"].synchronize()
            return sl["h_keep"], sl["h_num"]

        want_keep, want_num = self._nms(self.boxes, self.scores)  # the device-resident call on the same frames
        want_keep, want_num = want_keep.cpu(), want_num.cpu()
        for k in range(4):  # warm-up: buffers, streams, both symmetric buffers; and the loop returns what the plain call returns
            submit(k)
            hk, hn = result(k)
            assert torch.equal(hk, want_keep) and torch.equal(hn, want_num), "pipelined multi-GPU step != plain step"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            if k >= 2:
                hk, hn = result(k - 2)  # frees the slot step k is about to use
            submit(k)
        for k in range(max(0, steps - 2), steps):
            hk, hn = result(k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        self.d2h = hk.numel() * 8 + hn.numel() * 4
        return dt

    def strong_setup(self):
        """strong scaling: the config's own job size -- cfg2: 64 frames in total, cfg5: 256 frames x 10 classes in total --
        replicated on every rank and split by sharded.nms_batched_sharded, results all-gathered"""
        from lidardetection_b200 import synth

        torch = self.torch
        if self.which == "nms_cfg2":
            b, s = synth.cfg2(64, 4096, seed=synth.SEEDS["cfg2"])
        else:
            b5, s5 = synth.cfg5(256, 10, 1000, seed=synth.SEEDS["cfg5"])
            b, s = b5.reshape(-1, 1000, 7), s5.reshape(-1, 1000)
        self.s_boxes, self.s_scores = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
        return b.shape[0]

    def strong_step(self):
        from lidardetection_b200 import sharded

        return sharded.nms_batched_sharded(self.s_boxes, self.s_scores, self.thresh, max_keep=self.post)

    def result_for_gather(self, out):
        return list(out)

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        """dominant kernel = nms_lazy_kernel: rotated IoU of the kept boxes' rows (FP32-pipe bound).

        Work accounting (307 flops for a pair whose overlap is exactly 0, 818 otherwise):
          executed   the pairs the lazy kernel puts to the test (its own device counters): the rows of kept
                     boxes plus failed speculation -- this is the work `achieved` / `frac` are quoted on;
          reference  N(N-1)/2 pairs per problem, the mask the reference materialises; the same metric for the
                     mask + sweep formulation (LG_FLAG_NMS_FULL_MASK) is reported beside it.
        """
        torch = self.torch
        from lidardetection_b200 import _lib
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        P, N = self.scores.shape
        nz = tot = 0
        for f in range(0, P, max(1, P // 8)):  # p from a sample of problems (exact counts, our IoU == reference bits)
            order = self.scores[f].sort(0, descending=True)[1]
            ov = U.boxes_overlap_bev(self.boxes[f][order], self.boxes[f][order])
            nz += int((torch.triu(ov, 1) > 0).sum())
            tot += N * (N - 1) // 2
        p = nz / max(tot, 1)
        pairs_ref = P * N * (N - 1) // 2
        L = _lib.lib()
        order = self.scores.sort(1, descending=True)[1].contiguous()
        keep = torch.empty((P, N), dtype=torch.int64, device="cuda")
        num = torch.zeros((P,), dtype=torch.int32, device="cuda")
        ws = torch.empty(L.lg_nms_workspace_bytes(P, N), dtype=torch.uint8, device="cuda")
        st = _lib.stream_ptr(self.boxes.device)
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

        def phases(ph, flags):
            rc = L.lg_nms_batched_phases(_lib.ptr(self.boxes), _lib.ptr(order), None, P, N, self.thresh, _lib.ptr(ws), ws.numel(),
                                         _lib.ptr(keep), _lib.ptr(num), flags, st, 0, ph)
            _lib.check(rc, "lg_nms_batched_phases")

        def timed(plan, flags):
            phases(7, flags)
            times = {ph: [] for ph in plan}
            for _ in range(max(3, steps)):
                for ph in plan:
                    if ph != 1:
                        l2_flush(flush)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    phases(ph, flags)
                    e.record()
                    e.synchronize()
                    times[ph].append(s.elapsed_time(e))
            return {ph: float(np.mean(v)) for ph, v in times.items()}

        t_lazy = timed((4,), _lib.LG_FLAG_NONE)  # one kernel: records, candidate rows and the resolve
        off = L.lg_nms_stats_offset(P, N)
        tested, heavy, nonzero = (int(x) for x in ws[off:off + 24].view(torch.int64).cpu().tolist())
        t_full = timed((1, 2, 4), _lib.LG_FLAG_NMS_FULL_MASK)
        t_k = t_lazy[4] * 1e-3
        flops = 307.0 * (tested - nonzero) + 818.0 * nonzero
        achieved = flops / t_k / 1e12
        peak = fp32_peak or 74.4
        f_ref = 307.0 * (1 - p) + 818.0 * p
        return {"bound": "fp32", "kernel": "nms_lazy_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": "tools/peak_fp32.cu unrolled-FFMA microbenchmark measured live (MEASURED_PEAKS.json has no FP32 figure)"
                if fp32_peak else "theoretical 148 SM x 128 lanes x 2 x 1.965 GHz",
                "algorithmic": {"pairs_executed_per_launch": tested, "pairs_through_polygon_path": heavy, "pairs_nonzero": nonzero,
                                "flops_per_launch": flops, "accounting": "307 flops per zero-overlap pair, 818 otherwise (SURVEY 8d), on the pairs "
                                "the kernel evaluates: rows of kept boxes + failed speculation, counted on the device",
                                "reference_pairs_per_launch": pairs_ref, "reference_nonzero_fraction": p,
                                "reference_equivalent_tflops": pairs_ref * f_ref / t_k / 1e12},
                "kernel_ms": {"nms_lazy_kernel": t_lazy[4]},
                "full_mask_formulation": {"kernel": "nms_mask_kernel", "kernel_ms": {"nms_prep_kernel": t_full[1], "nms_mask_kernel": t_full[2],
                                                                                   "nms_sweep_kernel": t_full[4]},
                                          "achieved": pairs_ref * f_ref / (t_full[2] * 1e-3) / 1e12,
                                          "frac": pairs_ref * f_ref / (t_full[2] * 1e-3) / 1e12 / peak,
                                          "gpairs_per_s": pairs_ref / (t_full[2] * 1e-3) / 1e9}}

    def cpu_sample(self, pool):
        keep, dt = pool.nms_frame(self.boxes_np[0], self.scores_np[0], self.thresh)
        return 1.0 / dt, f"1 {'frame' if self.which == 'nms_cfg2' else 'problem'} of {self.boxes_np.shape[1]} boxes: boxes_iou_bev_cpu on the upper block-triangle in row blocks + host sweep (iou3d_nms.cpp:116-132)"


class IouWorkload(Workload):
    dtype = "f32"

    def __init__(self, torch, which, rank, world, light=False):  # light: device-resident timing only, no pinned host copies
        from lidardetection_b200 import synth

        self.torch, self.which = torch, which
        self.metric, self.unit = "rotated-IoU Gpairs/s", "Gpairs/s"
        if which == "iou_dense":
            a, b = synth.dense_overlap(16384, 16384, seed=synth.SEEDS["dense"] + rank)
            self.fn, self.name = "boxes_iou_bev", "dense-overlap microbench: boxes_iou_bev 16384 x 16384, all pairs overlapping, per GPU"
        elif which == "iou_cfg1":
            a, b = synth.cfg1(seed=synth.SEEDS["cfg1"] + rank)
            self.fn, self.name = "boxes_iou_bev", "PointPillars KITTI anchor-target IoU: boxes_iou_bev 321,408 anchors x 20 GT, per GPU"
        else:
            a, b = synth.cfg4(200_000)
            rows = 25_000  # one eighth of the 200k rows per GPU (20 GB of output), whatever N is
            r0 = (rank % 8) * rows
            a = a[r0:r0 + rows]
            self.fn, self.name = "boxes_iou3d_gpu", "Waymo-scale evaluation IoU: boxes_iou3d 25,000-row shard x 200,000 boxes per GPU (200k x 200k over 8 shards)"
        self.a_np, self.b_np = a, b
        self.a, self.b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        if not light:
            self.h_a, self.h_b = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
        self.pairs = a.shape[0] * b.shape[0]
        self.units = self.pairs / 1e9
        self.out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device="cuda")
        # iou_flat_kernel alone (M <= 64); else prep_kernel + density_probe_kernel + the two iou_strip_kernel variants (one of them exits at
        # once); from 2^26 pairs on also iou_sweep_kernel + iou_pairs_kernel (two-phase sweep; the sparse strip kernel then only mops up)
        self.launches_per_step = 1 if b.shape[0] <= 64 else (6 if self.pairs >= (1 << 26) else 4)
        self.h2d = (a.size + b.size) * 4
        self.e2e_d2h_full = self.pairs * 4 <= (1 << 30)
        self.d2h = self.pairs * 4 if self.e2e_d2h_full else a.shape[0] * 8
        self.h_out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32).pin_memory() if (self.e2e_d2h_full and not light) else None

    def _call(self, a, b):
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        name = {"boxes_iou_bev": "lg_boxes_iou_bev", "boxes_iou3d_gpu": "lg_boxes_iou3d"}[self.fn]
        return U._iou_call(name, a, b, out=self.out)  # same C-ABI call as U.<fn>, output buffer reused

    def step(self):
        return self._call(self.a, self.b)

    def e2e_step(self):
        a, b = self.h_a.cuda(non_blocking=True), self.h_b.cuda(non_blocking=True)
        out = self._call(a, b)
        if self.e2e_d2h_full:
            self.h_out.copy_(out, non_blocking=False)
            return self.h_out
        # 20 GB result: what the evaluation consumes is the per-row best match (max, argmax), read back
        mx, am = out.max(1)
        return mx.cpu(), am.cpu()

    def result_for_gather(self, out):
        return []  # row blocks stay sharded (160 GB is never gathered)

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        torch = self.torch
        from lidardetection_b200 import _lib
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        n, m = self.a.shape[0], self.b.shape[0]
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            l2_flush(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        if self.which == "iou_dense":
            nz = float((self.out[:2048] > 0).float().mean())
            flops = self.pairs * (307.0 * (1 - nz) + 818.0 * nz)
            ach, peak = flops / t / 1e12, (fp32_peak or 74.4)
            return {"bound": "fp32", "kernel": "iou_strip_kernel (+prep_kernel, <1%)", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "peak_source": "tools/peak_fp32.cu FFMA microbenchmark, measured live",
                    "algorithmic": {"pairs_per_launch": self.pairs, "nonzero_fraction": nz, "flops_per_pair": 307.0 * (1 - nz) + 818.0 * nz}}
        byts = 4.0 * n * m + 28.0 * (n + m)
        ach = byts / t / 1e9
        kern = ("iou_flat_kernel" if m <= 64 else
                "iou_sweep_kernel (zeros + survivor list, ~80 % of the step) + iou_pairs_kernel (polygon path on the list) + prep / probe"
                if self.pairs >= (1 << 26) else "iou_strip_kernel (+prep_kernel, density_probe_kernel)")
        r = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
             "traffic": None, "peak_source": hbm_src,
             "algorithmic": {"bytes_per_launch": byts, "formula": "4*N*M + 28*(N+M)", "time": "the whole call (every launch of the step), CUDA events"}}
        if self.which == "iou_cfg1":
            r["note"] = ("launch / tail bound, not bandwidth bound: one 35 us launch writes 26 MB, which is still in the 126 MB L2 when the "
                         "kernel ends (DRAM traffic under ncu is a third of the algorithmic bytes); the fraction measures launch latency")
        return r

    def cpu_sample(self, pool):
        rows = {"iou_dense": 2048, "iou_cfg1": 321408, "iou_cfg4": 2000}[self.which]
        cols = {"iou_dense": 2048}.get(self.which, self.b_np.shape[0])
        a, b = self.a_np[:rows], self.b_np[:cols]
        dt = pool.iou_matrix(a, b)
        return a.shape[0] * b.shape[0] / dt / 1e9, f"boxes_iou_bev_cpu on a {a.shape[0]} x {b.shape[0]} slice in row blocks (rows are independent: linear extrapolation)"


class PibWorkload(Workload):
    dtype = "f32"

    def __init__(self, torch, rank, light=False):
        from lidardetection_b200 import synth

        self.torch = torch
        B = 4096
        base_p, base_r = synth.cfg3(n_frames=64, seed=synth.SEEDS["cfg3"] + rank)
        rep = B // 64
        self.pts_np, self.rois_np = np.tile(base_p, (rep, 1, 1)), np.tile(base_r, (rep, 1, 1))
        self.name = "PV-RCNN KITTI points_in_boxes_gpu: 16,384 points x 100 ROIs per frame, 4096 frames per GPU (64 distinct, tiled)"
        self.metric, self.unit = "points-in-boxes frames/s (16,384 pts x 100 ROIs)", "frames/s"
        self.units = B
        self.pts, self.rois = torch.from_numpy(self.pts_np).cuda(), torch.from_numpy(self.rois_np).cuda()
        if not light:
            self.h_pts, self.h_rois = torch.from_numpy(self.pts_np).pin_memory(), torch.from_numpy(self.rois_np).pin_memory()
        self.launches_per_step = 1
        self.h2d = (self.pts_np.size + self.rois_np.size) * 4
        self.d2h = B * 16384 * 4

    def step(self):
        from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU

        return PU.points_in_boxes_gpu(self.pts, self.rois)

    def e2e_step(self):
        from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU

        return PU.points_in_boxes_gpu(self.h_pts.cuda(non_blocking=True), self.h_rois.cuda(non_blocking=True)).cpu()

    def result_for_gather(self, out):
        return []

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        torch = self.torch
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        byts = self.units * (12.0 * 16384 + 28.0 * 100 + 4.0 * 16384)
        ach = byts / t / 1e9
        return {"bound": "hbm", "kernel": "pib_grid_kernel", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "peak_source": hbm_src, "algorithmic": {"bytes_per_launch": byts, "formula": "B*(12*M + 28*T + 4*M)", "bytes_per_frame": 264944}}

    def cpu_sample(self, pool):
        n = 4 * pool.workers
        dt = pool.points_mask(self.pts_np[:n], self.rois_np[:n])
        return n / dt, f"points_in_boxes_cpu on {n} frames of 16,384 points x 100 boxes (one frame per task)"


class RoiPoolWorkload(Workload):
    """SURVEY 8f-3.  roiaware_partA2: the two RoIAwarePool3d calls per frame of PartA2Head.roiaware_pool (partA2_head.py:131-143;
    PartA2.yaml: 128 ROIs, 12^3 voxels, 128 points per voxel): avg over 4 part-location channels + max over 128 RPN channels.
    roipoint_pointrcnn: RoIPointPool3d of PointRCNNHead.roipool3d_gpu (pointrcnn_head.py:117-121; 128 ROIs, 512 samples,
    128 + 2 feature channels, batch 4)."""
    dtype = "f32"
    cpu_kind = "port"

    def __init__(self, torch, which, rank):
        from lidardetection_b200 import synth
        from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU
        from lidardetection_b200.ops.roipoint_pool3d import roipoint_pool3d_utils as RU

        self.torch, self.which = torch, which
        B, M, N = 4, 16384, 128
        self.units = B
        frames = [synth.pool_case(M, N, 130, seed=9000 + 16 * rank + f) for f in range(B)]
        self.np_pts = np.stack([f[0] for f in frames])
        self.np_rois = np.stack([f[1] for f in frames])
        feat = np.stack([f[2] for f in frames])
        if which == "roiaware_partA2":
            self.np_feats = [np.ascontiguousarray(feat[..., :4]), np.ascontiguousarray(feat[..., 2:130])]
            self.layer = PU.RoIAwarePool3d(out_size=12, max_pts_each_voxel=128)
            V = 12 ** 3
            self.name = "Part-A2 KITTI RoI-aware pooling: 16,384 points x 128 ROIs, 12^3 voxels x 128 points, avg (C=4) + max (C=128) call per frame, 4 frames per GPU"
            self.metric, self.unit = "RoI-aware pooling frames/s (128 ROIs, 12^3 voxels)", "frames/s"
            # algorithmic bytes per frame: inputs once per call; outputs: voxel lists twice (one per call), pooled, argmax (max call)
            self.bytes_per_frame = 2 * (28.0 * N + 12.0 * M) + 4.0 * M * (4 + 128) + 2 * 4.0 * N * V * 128 + 4.0 * N * V * (4 + 128 + 128)
            self.formula = "2*(28N + 12M) + 4M(4+128) + 2*4*N*V*max_pts + 4*N*V*(4 + 128 + 128)"
            self.kernel = "memset + roiaware_collect_kernel + roiaware_pool_kernel, both calls"
            self.launches_per_step = 4 * B
            self.d2h = B * 4 * N * V * (4 + 128)
        else:
            self.np_feats = [feat]
            self.layer = RU.RoIPointPool3d(num_sampled_points=512, pool_extra_width=(0.0, 0.0, 0.0))
            self.name = "PointRCNN KITTI RoI point pooling: 16,384 points x 128 ROIs x 512 samples, 130 channels, batch 4 in one call"
            self.metric, self.unit = "RoI point pooling frames/s (128 ROIs x 512 samples)", "frames/s"
            self.bytes_per_frame = 28.0 * N + 12.0 * M + 4.0 * M * 130 + 4.0 * N * 512 * 133 + 4.0 * N
            self.formula = "28N + 12M + 4*M*C + 4*N*S*(3+C) + 4N"
            self.kernel = "roipoint_pool_kernel"
            self.launches_per_step = 1
            self.d2h = B * (4 * N * 512 * 133 + 4 * N)
        cu = lambda x: torch.from_numpy(x).cuda()  # noqa: E731
        pin = lambda x: torch.from_numpy(x).pin_memory()  # noqa: E731
        self.pts, self.rois, self.feats = cu(self.np_pts), cu(self.np_rois), [cu(f) for f in self.np_feats]
        self.h_pts, self.h_rois, self.h_feats = pin(self.np_pts), pin(self.np_rois), [pin(f) for f in self.np_feats]
        self.h2d = 4 * (self.np_pts.size + self.np_rois.size + sum(f.size for f in self.np_feats))

    def _run(self, pts, rois, feats):
        if self.which == "roiaware_partA2":
            outs = []
            for b in range(pts.shape[0]):  # the reference's own per-frame loop (partA2_head.py:131)
                outs.append(self.layer(rois[b], pts[b], feats[0][b], pool_method="avg"))
                outs.append(self.layer(rois[b], pts[b], feats[1][b], pool_method="max"))
            return outs
        return list(self.layer(pts, feats[0], rois))

    def step(self):
        return self._run(self.pts, self.rois, self.feats)

    def e2e_step(self):
        outs = self._run(self.h_pts.cuda(non_blocking=True), self.h_rois.cuda(non_blocking=True), [f.cuda(non_blocking=True) for f in self.h_feats])
        return [o.cpu() for o in outs]

    def result_for_gather(self, out):
        return []

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        torch = self.torch
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        byts = self.units * self.bytes_per_frame
        ach = byts / t / 1e9
        return {"bound": "hbm", "kernel": self.kernel, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "peak_source": hbm_src, "algorithmic": {"bytes_per_step": byts, "formula": self.formula, "bytes_per_frame": self.bytes_per_frame,
                                                        "note": "timed over the whole step (all launches of the frames' calls), not one kernel"}}

    def cpu_sample(self, pool):
        n = pool.workers
        B = self.np_pts.shape[0]
        if self.which == "roiaware_partA2":
            jobs = [(self.np_pts[i % B], self.np_rois[i % B], self.np_feats[0][i % B], self.np_feats[1][i % B], 12, 128) for i in range(n)]
            dt = pool.roiaware_frames(jobs)
            what = "RoI-aware pooling"
        else:
            jobs = [(self.np_pts[i % B], self.np_rois[i % B], self.np_feats[0][i % B], 512) for i in range(n)]
            dt = pool.roipoint_frames(jobs)
            what = "RoI point pooling"
        return n / dt, f"C restatement of the reference CUDA kernels ({what}; the reference has no CPU build of it), {n} frames, one per core"


class KittiEvalWorkload(Workload):
    """the overlap half of one KITTI val evaluation -- calculate_iou_partly for metric 1 (bev) and metric 2 (3d)
    (eval.py:340-414 as called from eval_class, eval.py:473) on 3769 synthetic frames in 51 parts (28.1 M pairs per metric).
    The reference does, per part and metric, H2D -> numba.cuda kernel -> D2H (-> numba CPU pass for 3d); here each metric is
    ONE launch of lg_kitti_overlaps_parts over all parts."""
    dtype = "f32 geometry / f64 area sum and ratio"
    cpu_kind = "port"

    def __init__(self, torch, rank):
        from lidardetection_b200 import synth
        from lidardetection_b200.datasets.kitti.kitti_object_eval_python import eval as E

        self.torch, self.E = torch, E
        self.gts, self.dts = synth.kitti_eval_frames(3769, 5000 + rank)
        parts = E.get_split_parts(3769, 50)
        self.gc, self.dc, i = [], [], 0
        for n in parts:
            self.gc.append(sum(len(x) for x in self.gts[i:i + n]))
            self.dc.append(sum(len(x) for x in self.dts[i:i + n]))
            i += n
        self.parts = parts
        self.G, self.D = np.concatenate(self.gts), np.concatenate(self.dts)
        self.g, self.d = torch.from_numpy(self.G).cuda(), torch.from_numpy(self.D).cuda()
        mk = lambda fr: [{"name": np.zeros(len(f)), "location": f[:, 0:3], "dimensions": f[:, 3:6], "rotation_y": f[:, 6]} for f in fr]  # noqa: E731
        self.ga, self.da = mk(self.gts), mk(self.dts)
        self.pairs = int(sum(a * b for a, b in zip(self.gc, self.dc)))
        self.units = 2 * self.pairs / 1e9
        self.metric, self.unit = "KITTI-eval rotated overlap Gpairs/s (bev + 3d)", "Gpairs/s"
        self.name = (f"KITTI val evaluation overlaps: calculate_iou_partly metric 1 (bev) + metric 2 (3d), 3769 frames in {len(parts)} parts, "
                     f"{self.pairs} pairs per metric, per GPU")
        self.launches_per_step = 8  # per metric: kitti_tiles_kernel + 2 x kitti_prep_kernel + kitti_pair_kernel
        self.h2d = 2 * 56 * (len(self.G) + len(self.D))
        # calculate_iou_partly brings back only the non-zero entries (int64 index + float32 value), counted here once
        nnz = sum(int((o.view(torch.int32) != 0).sum()) for o in self.step())
        self.d2h = 12 * nnz + 16
        self.bytes_per_step = 2.0 * (4.0 * self.pairs + 56.0 * (len(self.G) + len(self.D)))

    def step(self):
        return [self.E.kitti_overlaps_parts_cuda(self.g, self.d, self.gc, self.dc, m)[0] for m in (1, 2)]

    def e2e_step(self):
        return [self.E.calculate_iou_partly(self.ga, self.da, m)[1] for m in (1, 2)]

    def result_for_gather(self, out):
        return []

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        torch = self.torch
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            l2_flush(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        ach = self.bytes_per_step / t / 1e9
        return {"bound": "hbm", "kernel": "kitti_pair_kernel (+2 x kitti_prep_kernel), both metrics", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src,
                "algorithmic": {"bytes_per_step": self.bytes_per_step, "formula": "2 metrics x (4 B x pairs + 56 B x boxes)",
                                "note": "99.5 % of the pairs are exactly 0 (culled): the output write is the algorithmic traffic; timed over the whole step"}}

    def cpu_sample(self, pool):
        n = min(pool.workers, len(self.parts))
        jobs, i = [], 0
        for p in range(n):
            jobs.append((np.concatenate(self.gts[i:i + self.parts[p]]), np.concatenate(self.dts[i:i + self.parts[p]])))
            i += self.parts[p]
        dt = pool.kitti_parts(jobs)
        pairs = sum(len(g) * len(d) for g, d in jobs)
        return 2 * pairs / dt / 1e9, (f"C restatement of the reference numba.cuda kernel + d3_box_overlap_kernel (the reference has no CPU build of the "
                                      f"rotated part), {n} evaluation parts (bev + 3d), one per core")


class PostProcWorkload(Workload):
    """the post-processing front end of SECOND KITTI (second.yaml:94-99): 70,400 candidates per frame,
    SCORE_THRESH 0.1 -> top 4096 -> rotated NMS (0.01) -> 500, 64 frames per GPU, one batched call."""
    dtype = "f32"

    def __init__(self, torch, rank):
        from lidardetection_b200 import synth

        self.torch = torch
        F, NC = 64, 70400
        bx, sc = synth.cfg2(n_frames=F, n_boxes=4096, seed=synth.SEEDS["cfg2"] + 1000 * rank)
        r = np.random.default_rng(17 + rank)
        rest = synth.gt_boxes(F * (NC - 4096), 9 + rank).reshape(F, -1, 7)
        low = r.uniform(0.0, 0.099, (F, NC - 4096)).astype(np.float32)  # below SCORE_THRESH: background anchors
        perm = r.permutation(NC)
        self.boxes_np = np.ascontiguousarray(np.concatenate([bx, rest], 1)[:, perm])
        self.scores_np = np.ascontiguousarray(np.concatenate([sc, low], 1)[:, perm])
        self.cfg = {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.01, "NMS_PRE_MAXSIZE": 4096, "NMS_POST_MAXSIZE": 500}
        self.name = ("SECOND KITTI post-processing front end: 70,400 candidates/frame, score >= 0.1 -> top 4096 -> rotated NMS 0.01 -> 500, "
                     "64 frames per GPU, one batched call (model_nms_utils.class_agnostic_nms_batched)")
        self.metric, self.unit = "post-processing frames/s (70,400 candidates -> 4096 -> NMS -> 500)", "frames/s"
        self.units = F
        self.boxes, self.scores = torch.from_numpy(self.boxes_np).cuda(), torch.from_numpy(self.scores_np).cuda()
        self.h_boxes, self.h_scores = torch.from_numpy(self.boxes_np).pin_memory(), torch.from_numpy(self.scores_np).pin_memory()
        self.launches_per_step = 4  # select_topk_kernel, nms_prep_kernel, nms_lazy_kernel, select_finish_kernel
        self.h2d = (self.boxes_np.size + self.scores_np.size) * 4
        self.d2h = 0

    def step(self):
        from lidardetection_b200 import model_nms_utils as MU

        return MU.class_agnostic_nms_batched(self.scores, self.boxes, self.cfg, score_thresh=0.1)

    def e2e_step(self):
        from lidardetection_b200 import model_nms_utils as MU

        sel, num, sc = MU.class_agnostic_nms_batched(self.h_scores.cuda(non_blocking=True), self.h_boxes.cuda(non_blocking=True), self.cfg,
                                                     score_thresh=0.1)
        num_h = num.cpu()
        k = int(num_h.max())
        sel_h, sc_h = sel[:, :k].cpu(), sc[:, :k].cpu()
        self.d2h = num_h.numel() * 4 + sel_h.numel() * 8 + sc_h.numel() * 4
        return sel_h, sc_h, num_h

    def result_for_gather(self, out):
        return []

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        """the step is torch.topk over (64, 70400) + a gather + nms_prep_kernel + nms_lazy_kernel; the lazy kernel dominates and
        is the kernel of nms_cfg2's roofline (same 4096-box problems).  Reported here: the HBM view of the whole step --
        algorithmic bytes = scores + boxes read once + results."""
        torch = self.torch
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            l2_flush(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        byts = float(self.boxes_np.size * 4 + self.scores_np.size * 4 + self.units * 500 * 12)
        ach = byts / t / 1e9
        return {"bound": "hbm", "kernel": "select_topk_kernel + nms_prep_kernel + nms_lazy_kernel + select_finish_kernel (whole step)", "achieved": ach, "peak": hbm_peak,
                "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src,
                "algorithmic": {"bytes_per_launch": byts, "formula": "32 B x candidates + 12 B x 500 x frames", "step_ms": t * 1e3}}

    def cpu_sample(self, pool):
        # the reference has no CPU NMS; per frame the same baseline as nms_cfg2 on the 4096 boxes that pass the threshold
        m = self.scores_np[0] >= 0.1
        keep, dt = pool.nms_frame(self.boxes_np[0][m], self.scores_np[0][m], 0.01)
        return 1.0 / dt, "1 frame: the 4096 boxes above the score threshold, boxes_iou_bev_cpu on the upper block-triangle + host sweep"


class IouMaxWorkload(Workload):
    """Waymo-scale evaluation matching -- per-box best match (row and column max / argmax of boxes_iou3d) of a
    25,000-row shard against 200,000 boxes, the 20 GB block never materialised; at N > 1 the column maxima of the shards
    are combined with one NCCL all-reduce(MAX) of 200,000 packed keys (the only exchange step of the path)."""
    dtype = "f32"

    def __init__(self, torch, rank, world):
        from lidardetection_b200 import synth

        self.torch, self.world = torch, world
        a, b = synth.cfg4(200_000)
        rows = 25_000
        self.r0 = (rank % 8) * rows
        self.a_np, self.b_np = a[self.r0:self.r0 + rows], b
        self.a, self.b = torch.from_numpy(self.a_np).cuda(), torch.from_numpy(b).cuda()
        self.h_a, self.h_b = torch.from_numpy(self.a_np).pin_memory(), torch.from_numpy(b).pin_memory()
        self.pairs = rows * b.shape[0]
        self.units = self.pairs / 1e9
        self.metric, self.unit = "rotated-IoU best-match Gpairs/s", "Gpairs/s"
        self.name = ("Waymo-scale evaluation matching: boxes_iou_max (row + column max/argmax of boxes_iou3d) 25,000-row shard x 200,000 boxes "
                     "per GPU, matrix never materialised")
        self.launches_per_step = 5  # prep, key init, iou_strip_kernel<reduce>, 2 x key unpack
        self.h2d = (self.a_np.size + b.size) * 4
        self.d2h = rows * 12 + b.shape[0] * 12

    def _run(self, a, b):
        import torch.distributed as dist

        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        torch = self.torch
        rmax, rarg, cmax, carg = U.boxes_iou_max(a, b, kind="iou3d", rows=True, cols=True)
        if self.world > 1:
            key = (cmax.view(torch.int32).to(torch.int64) << 32) | (0xFFFFFFFF - (carg + self.r0))
            dist.all_reduce(key, op=dist.ReduceOp.MAX)
            cmax = (key >> 32).to(torch.int32).view(torch.float32)
            carg = 0xFFFFFFFF - (key & 0xFFFFFFFF)
        return rmax, rarg, cmax, carg

    def step(self):
        return self._run(self.a, self.b)

    def local_step(self):  # the same work without the all-reduce
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        return U.boxes_iou_max(self.a, self.b, kind="iou3d", rows=True, cols=True)

    def e2e_step(self):
        return tuple(x.cpu() for x in self._run(self.h_a.cuda(non_blocking=True), self.h_b.cuda(non_blocking=True)))

    def result_for_gather(self, out):
        return []

    def roofline(self, steps, hbm_peak, hbm_src, fp32_peak):
        """No matrix leaves the chip, so the bound is the FP32 / issue work of the pair tests themselves.  Algorithmic work per
        pair: 8 flops for the exact-zero cull test every pair needs (|ca - cb|^2 vs (ra + rb)^2), 818 for a pair whose overlap
        is non-zero; the non-zero count is exact (counted on the shard's row block through the matrix entry point)."""
        torch = self.torch
        from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U

        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.step()  # untimed: the caching allocator settles (the flush buffer above may have taken the block the workspace was using)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(3, steps)):
            l2_flush(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            U.boxes_iou_max(self.a, self.b, kind="iou3d", rows=True, cols=True)
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        t = float(np.median(ts)) * 1e-3
        nz = int((U.boxes_iou3d_gpu(self.a[:2000], self.b) > 0).sum()) * (self.a.shape[0] / 2000.0)
        flops = 8.0 * (self.pairs - nz) + 818.0 * nz
        ach, peak = flops / t / 1e12, (fp32_peak or 74.4)
        return {"bound": "fp32", "kernel": "iou_strip_kernel<reduce> (+prep, key init / unpack < 1%)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "peak_source": "tools/peak_fp32.cu FFMA microbenchmark, measured live",
                "algorithmic": {"pairs_per_launch": self.pairs, "nonzero_pairs_estimate": nz, "flops_per_launch": flops,
                                "accounting": "8 flops per culled pair (the circle test), 818 per pair with non-zero overlap", "kernel_ms": t * 1e3}}

    def cpu_sample(self, pool):
        a, b = self.a_np[:2000], self.b_np
        dt = pool.iou_matrix(a, b)
        return a.shape[0] * b.shape[0] / dt / 1e9, f"boxes_iou_bev_cpu on a {a.shape[0]} x {b.shape[0]} slice in row blocks (the reference has to build the matrix to take its max)"


# ------------------------------------------------------------------------------------------------
def timed_steps(torch, fn, steps, warmup, flush, world):
    """CUDA-event time of `steps` calls of fn (L2 flushed before each, outside the events), summed, max over ranks -> ms"""
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        l2_flush(flush)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ms += s.elapsed_time(e)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def run_secondary(torch, which, rank, world, hbm_peak, hbm_src, fp32_peak, flush):
    """the other halves of BASELINE.json's metric, measured in the same driver-run invocation: rotated-IoU Gpairs/s (dense =
    FP32 roofline, cfg4 shard = HBM roofline) and points-in-boxes frames/s.  Every rank runs its own shard / replica (no
    collective: SURVEY 8e, the matrices and index blocks stay with their owner); time = max over ranks."""
    wl = make_workload(torch, which, rank, world, light=True)
    steps = 5
    ms = timed_steps(torch, wl.step, steps, 3, flush, world)
    out = {"workload": wl.name, "metric": wl.metric, "unit": wl.unit, "value": wl.units * world * steps / (ms * 1e-3), "ms_per_step": ms / steps,
           "steps": steps, "warmup": 3, "gpu_launches_per_step": wl.launches_per_step, "dtype": wl.dtype}
    if rank == 0:
        r = wl.roofline(3, hbm_peak, hbm_src, fp32_peak)
        tr, src = ncu_traffic(which)
        if tr is not None:
            r["traffic"], r["traffic_source"] = tr, src
        out["roofline"] = r
    del wl
    torch.cuda.empty_cache()
    return out


def gpu_baseline(torch, wl):
    """the reference's own CUDA kernels (oracle/_ref: the unmodified sources compiled for sm_100a) on this GPU, outside every timed
    region: boxes_iou_bev_gpu on the dense 16384^2 microbenchmark (iou3d_nms_kernel.cu:251-265) and the per-frame nms_gpu of the
    default workload (iou3d_nms.cpp:90-136: kernel, blocking D2H of the mask, host sweep), plus OUR per-frame nms_gpu -- the
    reference's actual call pattern (detector3d_template.py:190-260) -- next to the batched call the headline times."""
    from lidardetection_b200 import synth
    from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
    from oracle import ref_loader as R

    res = {}
    frames = min(16, wl.boxes.shape[0])
    for _ in range(2):
        for f in range(frames):
            U.nms_gpu(wl.boxes[f], wl.scores[f], wl.thresh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(frames):
        U.nms_gpu(wl.boxes[f], wl.scores[f], wl.thresh)  # one frame per call, the keep count read back per call, as the reference API forces
    torch.cuda.synchronize()
    res["ours_per_frame_nms_gpu_ms"] = 1e3 * (time.perf_counter() - t0) / frames
    if not R.available():
        res["reference_cuda"] = "oracle/_ref did not travel with this snapshot"
        return res
    ref = R.iou3d_nms_cuda()
    a, b = synth.dense_overlap(16384, 16384, seed=synth.SEEDS["dense"])
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.zeros((16384, 16384), device="cuda")
    ref.boxes_iou_bev_gpu(ta, tb, out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(2):
        ref.boxes_iou_bev_gpu(ta, tb, out)
    e.record()
    e.synchronize()
    res["reference_boxes_iou_bev_gpu_dense_gpairs"] = 2 * 16384 * 16384 / (s.elapsed_time(e) * 1e-3) / 1e9
    del out, ta, tb
    order = wl.scores.sort(1, descending=True)[1]
    sorted_boxes = [wl.boxes[f][order[f]].contiguous() for f in range(frames)]
    keep = torch.LongTensor(wl.boxes.shape[1])
    ref.nms_gpu(sorted_boxes[0], keep, wl.thresh)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(frames):
        ref.nms_gpu(sorted_boxes[f], keep, wl.thresh)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / frames
    res["reference_nms_gpu_per_frame_ms"] = 1e3 * dt
    res["reference_nms_gpu_frames_per_s"] = 1.0 / dt
    res["what"] = ("reference CUDA kernels recompiled unmodified for sm_100a (oracle/_ref), same GPU, outside the timed regions; nms_gpu timed on "
                   f"{frames} pre-sorted frames of {wl.boxes.shape[1]} boxes, one call per frame (its API has no batch dimension)")
    return res


def make_workload(torch, which, rank, world, light=False):
    if which == "post_cfg2":
        return PostProcWorkload(torch, rank)
    if which == "iou_max_cfg4":
        return IouMaxWorkload(torch, rank, world)
    if which in ("nms_cfg2", "nms_cfg5"):
        return NmsWorkload(torch, which, rank, world)
    if which in ("iou_dense", "iou_cfg1", "iou_cfg4"):
        return IouWorkload(torch, which, rank, world, light)
    if which == "pib_cfg3":
        return PibWorkload(torch, rank, light)
    if which in ("roiaware_partA2", "roipoint_pointrcnn"):
        return RoiPoolWorkload(torch, which, rank)
    if which == "kitti_eval":
        return KittiEvalWorkload(torch, rank)
    raise SystemExit(f"unknown workload {which}")


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuPool

    pool = CpuPool(prefer_reference=True)
    kind_override = None

    from lidardetection_b200 import synth

    # the two "next row" workloads time the reference CPU path of the row they extend (it has nothing else to offer there)
    which = {"post_cfg2": "nms_cfg2", "iou_max_cfg4": "iou_cfg4"}.get(args.workload, args.workload)
    if which in ("nms_cfg2", "nms_cfg5"):
        if which == "nms_cfg2":
            b, s = synth.cfg2(2, 4096)
            thr, name, metric, unit, n = 0.01, "SECOND KITTI post-processing: rotated nms_gpu 4096 boxes/frame, thresh 0.01, 64 frames per GPU", "rotated NMS frames/s (4096 boxes/frame)", "frames/s", 4096
        else:
            b5, s5 = synth.cfg5(1, 10, 1000)
            b, s = b5.reshape(-1, 1000, 7), s5.reshape(-1, 1000)
            thr, name, metric, unit, n = 0.2, "NuScenes CBGS multi-head NMS: 10 classes x 1000 boxes x 256 frames per GPU, thresh 0.2", "rotated NMS problems/s (1000 boxes/problem)", "problems/s", 1000
        sample = f"1 problem of {n} boxes per step: boxes_iou_bev_cpu on the upper block-triangle in row blocks + host sweep"

        def one(i):
            return pool.nms_frame(b[i % b.shape[0]], s[i % b.shape[0]], thr)[1], 1.0
    elif which in ("iou_dense", "iou_cfg1", "iou_cfg4"):
        metric, unit = "rotated-IoU Gpairs/s", "Gpairs/s"
        if which == "iou_dense":
            a, bb = synth.dense_overlap(2048, 2048)
            name = "dense-overlap microbench: boxes_iou_bev 16384 x 16384, all pairs overlapping, per GPU"
        elif which == "iou_cfg1":
            a, bb = synth.cfg1()
            name = "PointPillars KITTI anchor-target IoU: boxes_iou_bev 321,408 anchors x 20 GT, per GPU"
        else:
            a, bb = synth.cfg4(200_000)
            a = a[:2000]
            name = "Waymo-scale evaluation IoU: boxes_iou3d 25,000-row shard x 200,000 boxes per GPU (200k x 200k over 8 shards)"
        sample = f"boxes_iou_bev_cpu {a.shape[0]} x {bb.shape[0]} per step in row blocks"

        def one(i):
            return pool.iou_matrix(a, bb), a.shape[0] * bb.shape[0] / 1e9
    elif which in ("roiaware_partA2", "roipoint_pointrcnn"):
        # the reference has no CPU build of these functions: the C restatement of its CUDA kernels (kind "port")
        frames = [synth.pool_case(16384, 128, 130, seed=9000 + f) for f in range(4)]
        kind_override = "port"
        if which == "roiaware_partA2":
            jobs = [(f[0], f[1], np.ascontiguousarray(f[2][:, :4]), np.ascontiguousarray(f[2][:, 2:130]), 12, 128) for f in frames]
            metric, unit, name = "RoI-aware pooling frames/s (128 ROIs, 12^3 voxels)", "frames/s", "Part-A2 KITTI RoI-aware pooling: 16,384 points x 128 ROIs, 12^3 voxels x 128 points, avg (C=4) + max (C=128) call per frame"
            run = pool.roiaware_frames
        else:
            jobs = [(f[0], f[1], f[2], 512) for f in frames]
            metric, unit, name = "RoI point pooling frames/s (128 ROIs x 512 samples)", "frames/s", "PointRCNN KITTI RoI point pooling: 16,384 points x 128 ROIs x 512 samples, 130 channels"
            run = pool.roipoint_frames
        jobs = [jobs[i % 4] for i in range(pool.workers)]
        sample = f"C restatement of the reference CUDA kernels, {len(jobs)} frames per step, one per core"

        def one(i):
            return run(jobs), float(len(jobs))
    elif which == "kitti_eval":
        from lidardetection_b200.datasets.kitti.kitti_object_eval_python.eval import get_split_parts

        kind_override = "port"  # the reference's rotated overlap exists only as a numba.cuda kernel
        gts, dts = synth.kitti_eval_frames(3769, 5000)
        parts = get_split_parts(3769, 50)
        jobs, i0 = [], 0
        for p in range(min(pool.workers, len(parts))):
            jobs.append((np.concatenate(gts[i0:i0 + parts[p]]), np.concatenate(dts[i0:i0 + parts[p]])))
            i0 += parts[p]
        metric, unit = "KITTI-eval rotated overlap Gpairs/s (bev + 3d)", "Gpairs/s"
        name = "KITTI val evaluation overlaps: calculate_iou_partly metric 1 (bev) + metric 2 (3d), 3769 frames in 51 parts"
        npairs = sum(len(g) * len(d) for g, d in jobs)
        sample = f"C restatement of the reference numba.cuda kernel + d3_box_overlap_kernel, {len(jobs)} evaluation parts (bev + 3d) per step, one per core"

        def one(i):
            return pool.kitti_parts(jobs), 2 * npairs / 1e9
    else:
        p, r = synth.cfg3(n_frames=4 * pool.workers)
        metric, unit = "points-in-boxes frames/s (16,384 pts x 100 ROIs)", "frames/s"
        name = "PV-RCNN KITTI points_in_boxes_gpu: 16,384 points x 100 ROIs per frame, 4096 frames per GPU (64 distinct, tiled)"
        sample = f"points_in_boxes_cpu on {p.shape[0]} frames per step"

        def one(i):
            return pool.points_mask(p, r), float(p.shape[0])
    for i in range(args.warmup):
        one(i)
    tot_t = tot_u = 0.0
    for i in range(args.steps):
        dt, u = one(i)
        tot_t += dt
        tot_u += u
    pool.close()
    val = tot_u / tot_t
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded, SURVEY.md 8d shapes)", "config": workload_config(name),
            "cpu_baseline": {"value": val, "unit": unit, "cores": pool.workers, "kind": kind_override or pool.kind, "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly ONE line, the JSON record: libraries that print there (NCCL's version banner does) are sent to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nms_cfg2", choices=["nms_cfg2", "nms_cfg5", "iou_dense", "iou_cfg1", "iou_cfg4", "pib_cfg3", "post_cfg2", "iou_max_cfg4", "roiaware_partA2", "roipoint_pointrcnn", "kitti_eval"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary metrics and the GPU baseline of the default workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # CPU baseline first: its worker pool forks, which must happen before CUDA is initialised
    cpu_pool = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuPool

        cpu_pool = CpuPool(prefer_reference=True)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime

        # a mismatched collective must fail in minutes, not hang the box for the watchdog's default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=90))
    from lidardetection_b200 import _lib

    _lib.check(_lib.lib().lg_check_device(), "lg_check_device")
    wl = make_workload(torch, args.workload, rank, world)

    cpu_baseline = None
    if cpu_pool is not None:
        v, sample = wl.cpu_sample(cpu_pool)
        cpu_baseline = {"value": v, "unit": wl.unit, "cores": cpu_pool.workers, "kind": getattr(wl, "cpu_kind", cpu_pool.kind), "sample": sample}
        cpu_pool.close()
        log("cpu_baseline", cpu_baseline)

    hbm_peak, hbm_src, _ = measured_peaks()
    fp32_peak = fp32_peak_tflops(torch)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        wl.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        l2_flush(flush)  # L2 flush between timed iterations (outside the events)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        wl.step()
        e.record()
        e.synchronize()
        dev_ms += s.elapsed_time(e)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    # keep the GPU under load a little longer if the region was too short for a clock sample
    t_end = time.perf_counter() + max(0.0, 0.6 - t_wall)
    while time.perf_counter() < t_end:  # a wall-clock loop: every rank runs its own number of trips, so nothing collective in it
        getattr(wl, "local_step", wl.step)()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the public API from pinned host memory --------------------------------
    for _ in range(2):
        wl.e2e_step()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gather_out = wl.e2e_step()
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
        del gather_out
    barrier()

    # N = 1, NMS workloads: the same steps through the serving loop of the public API (four steps in flight); the per-call figure
    # above stays in the line as e2e_sync
    e2e_sync_s, e2e_depth = e2e_s, 1
    if hasattr(wl, "e2e_pipelined"):
        e2e_depth = 4 if world == 1 else 2
        barrier()
        e2e_s = wl.e2e_pipelined(args.steps, e2e_depth)
        barrier()
    tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(tt[0]), float(tt[1])

    # ---- strong scaling (N > 1, NMS workloads): the config's own total job split over the ranks, keep lists all-gathered
    strong = None
    if world > 1 and hasattr(wl, "strong_setup"):
        total = wl.strong_setup()
        ms = timed_steps(torch, wl.strong_step, args.steps, args.warmup, flush, world)
        strong = {"value": total * args.steps / (ms * 1e-3), "unit": wl.unit, "ms_per_step": ms / args.steps, "units_total": total,
                  "what": f"{total} problems in total, contiguous blocks of ceil({total} / {world}) per rank (sharded.nms_batched_sharded), the packed "
                          "(block, 1 + NMS_POST_MAXSIZE) int64 results gathered on every rank inside the timed region (fused peer-memory gather)"}

    roofline = wl.roofline(args.steps, hbm_peak, hbm_src, fp32_peak) if rank == 0 else None
    if roofline is not None:
        tr, src = ncu_traffic(args.workload)
        if tr is not None:
            roofline["traffic"], roofline["traffic_source"] = tr, src
    gbase = None
    if rank == 0 and world == 1 and isinstance(wl, NmsWorkload) and not args.no_secondary:
        gbase = gpu_baseline(torch, wl)
    # ---- the other halves of the metric (default workload only): rotated-IoU Gpairs/s and points-in-boxes frames/s
    secondary = None
    if args.workload == "nms_cfg2" and not args.no_secondary:
        del wl.boxes, wl.scores
        torch.cuda.empty_cache()
        secondary = {w: run_secondary(torch, w, rank, world, hbm_peak, hbm_src, fp32_peak, flush) for w in ("iou_dense", "iou_cfg4", "pib_cfg3")}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    total_units = wl.units * world
    line = {
        "metric": wl.metric, "value": total_units * args.steps / (dev_ms * 1e-3), "unit": wl.unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl.dtype, "data": "synthetic (seeded, SURVEY.md 8d shapes; no datasets offline)",
        # `config` is the workload and nothing else, identical in both arms (the driver compares them); how THIS arm measures it is
        # in `measurement`
        "config": workload_config(wl.name),
        "measurement": {"units_per_step_per_gpu": wl.units,
                        "timing": "CUDA events per step on the launching stream, summed over steps, max over ranks",
                        "e2e": (("HostNmsPipeline (public API)" if world == 1 else "the same serving loop around sharded.nms_batched_sharded (upload on one stream; kernels + gather + download of the gathered lists on another)") +
                                ": every step uploads its inputs from pinned host memory and downloads its result; "
                                f"{e2e_depth} steps in flight (upload of step k + 1 overlaps the kernels of step k); wall clock over all steps"
                                if e2e_depth > 1 else "one blocking call of the public API per step from pinned host memory, result read back; wall clock"),
                        "multi_gpu": (getattr(wl, "multi_gpu_note", None) or
                                      "independent problems per rank, no data-path collective (results stay on the owning rank, as in the "
                                      "reference's DDP evaluation); NCCL only for the barrier and the max-over-ranks of the timings")
                        if world > 1 else "single GPU"},
        "e2e": {"value": total_units * args.steps / e2e_s, "unit": wl.unit, "h2d_bytes_per_step": int(wl.h2d), "d2h_bytes_per_step": int(wl.d2h)},
        "e2e_sync": {"value": total_units * args.steps / e2e_sync_s, "unit": wl.unit,
                     "what": "one blocking call per step: upload, kernels, download, synchronise -- nothing of step k + 1 starts before step k's result is on the host"},
        "gpu_launches": wl.launches_per_step * args.steps,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "strong": strong, "secondary": secondary, "gpu_baseline": gbase,
        "fp32_peak_tflops_measured": fp32_peak, "hbm_peak_gbs": hbm_peak,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
