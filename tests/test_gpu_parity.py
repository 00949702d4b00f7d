"""GPU parity tests proper: liblidargeom (through the Python drop-in modules -> ctypes -> C ABI) against
  (1) the oracle restatement (oracle/lg_oracle.c, CUDA flavor) on the same seeded inputs,
  (2) the committed golden vectors of the reference CUDA kernels (tests/golden/golden_gpu.npz),
  (3) when oracle/_ref travelled with the snapshot, the reference CUDA kernels run live on this GPU.
Tolerances (north_star): IoU <= 1e-5 absolute; NMS keep indices and point-in-box results bit-exact, except
documented ties where the deciding comparison is within 1e-6 of its threshold (none occur on these inputs:
the tests assert exact equality and would report the deciding value otherwise).
"""
import os

import numpy as np
import pytest
import torch

from lidardetection_b200 import _lib, synth
from lidardetection_b200.ops.iou3d_nms import iou3d_nms_utils as U
from lidardetection_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as PU
from oracle import lg_oracle as O
from oracle import ref_loader as R

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
IOU_ATOL = 1e-5
IOU_SETS = ["kat", "kat_t", "kat_sq", "car35", "ped70", "mix150", "dense", "cfg3iou", "cfg1sub"]


def dev():
    return torch.device("cuda:0")


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev())


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def assert_iou_close(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert d.max() <= IOU_ATOL, f"{what}: max |diff| {d.max():.3g} > {IOU_ATOL}"
    # exact zeros must stay exact zeros (database_sampler.py:220 tests == 0)
    assert np.array_equal(got == 0, want == 0), f"{what}: zero pattern differs"
    return int((bits(got) != bits(want)).sum())


def explain_nms_mismatch(boxes_sorted, got_pos, want_pos, thresh, normal):
    """documented-tie rule: find the first differing decision and recompute its deciding IoU with the oracle"""
    for k in range(min(len(got_pos), len(want_pos))):
        if got_pos[k] != want_pos[k]:
            j = min(got_pos[k], want_pos[k])
            kept = want_pos[:k]
            vals = [O.iou_normal_pair(boxes_sorted[i], boxes_sorted[j]) if normal else O.iou_bev_pair(boxes_sorted[i], boxes_sorted[j])
                    for i in kept]
            near = [v for v in vals if abs(v - thresh) < 1e-6]
            return f"first difference at output {k} (box {j}); deciding IoUs within 1e-6 of {thresh}: {near}"
    return "length differs"


@pytest.fixture(scope="module")
def ggpu():
    p = os.path.join(HERE, "golden", "golden_gpu.npz")
    if not os.path.exists(p):
        pytest.skip("golden_gpu.npz not generated yet")
    return np.load(p)


def test_library_sees_a_blackwell_device():
    assert _lib.lib().lg_check_device() == 0, _lib.lib().lg_last_error_string()


# ------------------------------------------------------------------------------------------ IoU
@pytest.mark.parametrize("seed,centre,pri,n,m", [
    (1, (35, 17.5), synth.KITTI_PRIORS[:1], 300, 257),
    (2, (70, 35), synth.KITTI_PRIORS[1:2], 300, 129),   # pedestrians at range: the FP32-conditioning stress case
    (3, (150, 75), synth.KITTI_PRIORS, 200, 300),
    (4, (-75, 75), synth.WAYMO_PRIORS, 130, 33),
    (5, (0.2, -0.1), synth.KITTI_PRIORS, 64, 20),
])
def test_iou_family_vs_oracle(seed, centre, pri, n, m):
    a, b = synth.clustered_pairs(n, m, seed, centre, pri)
    ta, tb = cu(a), cu(b)
    nb = assert_iou_close(U.boxes_iou_bev(ta, tb).cpu().numpy(), O.boxes_iou_bev(a, b, O.FLAVOR_CUDA), "iou_bev")
    no = assert_iou_close(U.boxes_overlap_bev(ta, tb).cpu().numpy(), O.boxes_overlap_bev(a, b, O.FLAVOR_CUDA), "overlap_bev")
    n3 = assert_iou_close(U.boxes_iou3d_gpu(ta, tb).cpu().numpy(), O.boxes_iou3d(a, b, O.FLAVOR_CUDA), "iou3d")
    # the arithmetic contract is mirrored exactly; only vertex-order ties may move last bits
    assert max(nb, no, n3) <= max(1, n * m // 2000)
    # strict (CPU-build) arithmetic vs the CPU flavor: un-contracted FP32 and glibc's sinf / cosf restated on the device
    # (lg_trig.cuh) -- the same bits, except where a vertex order hangs on the last bit of atan2f (libdevice here, glibc there)
    strict = U._iou_call("lg_boxes_iou_bev", ta, tb, flags=_lib.LG_FLAG_STRICT_FP32).cpu().numpy()
    ns = assert_iou_close(strict, O.boxes_iou_bev(a, b, O.FLAVOR_CPU), "strict iou_bev vs CPU flavor")
    assert ns <= max(1, n * m // 2000)


@pytest.mark.parametrize("name", IOU_SETS)
def test_iou_family_vs_golden_reference_cuda(ggpu, name):
    ta, tb = cu(ggpu[f"iou_{name}_a"]), cu(ggpu[f"iou_{name}_b"])
    for key, fn in (("bev", U.boxes_iou_bev), ("overlap", U.boxes_overlap_bev), ("iou3d", U.boxes_iou3d_gpu)):
        assert_iou_close(fn(ta, tb).cpu().numpy(), ggpu[f"iou_{name}_{key}"], f"{name}/{key}")


def test_iou_shapes_strides_and_empties():
    a, b = synth.clustered_pairs(70, 45, 7)
    ta, tb = cu(a), cu(b)
    want = O.boxes_iou_bev(a, b, O.FLAVOR_CUDA)
    # callers pass boxes[:, 0:7] slices of wider tensors (non-contiguous)
    wide_a = torch.cat([ta, torch.randn(70, 3, device=dev())], 1)
    wide_b = torch.cat([tb, torch.randn(45, 2, device=dev())], 1)
    assert_iou_close(U.boxes_iou_bev(wide_a[:, 0:7], wide_b[:, 0:7]).cpu().numpy(), want, "sliced")
    for n, m in ((0, 5), (5, 0), (0, 0)):
        out = U.boxes_iou3d_gpu(ta[:n], tb[:m])
        assert out.shape == (n, m) and out.dtype == torch.float32 and out.is_cuda
    # every tile-width variant and ragged edges: m = 1, 31, 32, 33, 64, 65, 127, 128, 129
    big_a, big_b = synth.clustered_pairs(140, 129, 8)
    wantb = O.boxes_iou_bev(big_a, big_b, O.FLAVOR_CUDA)
    for m in (1, 31, 32, 33, 64, 65, 127, 128, 129):
        got = U.boxes_iou_bev(cu(big_a), cu(big_b[:m])).cpu().numpy()
        assert_iou_close(got, wantb[:, :m], f"m={m}")
    # unaligned output pitch goes through the scalar store path
    out = torch.empty((140, 131), device=dev())[:, :129]
    got = U._iou_call("lg_boxes_iou_bev", cu(big_a), cu(big_b), out=out).cpu().numpy()
    assert_iou_close(got, wantb, "pitched")


def test_iou_cfg1_shape_sampled_against_oracle():
    """PointPillars anchors x GT at full size (321,408 x 20); the oracle checks a row sample."""
    a, gt = synth.cfg1()
    got = U.boxes_iou_bev(cu(a), cu(gt)).cpu().numpy()
    assert got.shape == (321408, 20)
    rows = np.unique(np.concatenate([np.arange(0, 321408, 53), np.nonzero(got.max(1) > 0)[0][:4000]]))
    assert_iou_close(got[rows], O.boxes_iou_bev(a[rows], gt, O.FLAVOR_CUDA), "cfg1 sample")
    frac = (got > 0).mean()
    assert 0.0005 < frac < 0.02


# ------------------------------------------------------------------------------------------ NMS
@pytest.mark.parametrize("n,thresh", [(1, 0.1), (63, 0.1), (64, 0.3), (65, 0.01), (300, 0.7), (1024, 0.01), (1000, 0.2), (4096, 0.01)])
@pytest.mark.parametrize("normal", [False, True])
def test_nms_vs_oracle(n, thresh, normal):
    boxes, scores = synth.nms_frames(1, n, seed=100 + n)
    tb, ts = cu(boxes[0]), cu(scores[0])
    fn = U.nms_normal_gpu if normal else U.nms_gpu
    keep, none = fn(tb, ts, thresh, NMS_TYPE="nms_gpu", NMS_THRESH=thresh)  # the whole NMS config is splatted in
    assert none is None and keep.dtype == torch.int64 and keep.is_cuda and keep.is_contiguous()
    order = ts.sort(0, descending=True)[1].cpu().numpy()
    want = O.nms(boxes[0], scores[0], thresh, normal=normal, flavor=O.FLAVOR_CUDA, order=order)
    got = keep.cpu().numpy()
    if not np.array_equal(got, want):
        inv = np.empty(n, np.int64)
        inv[order] = np.arange(n)
        pytest.fail(explain_nms_mismatch(boxes[0][order], inv[got], inv[want], thresh, normal))


@pytest.mark.parametrize("n,thresh,seed", [(4096, 0.01, 1), (4096, 0.5, 2), (1000, 0.2, 3), (777, 0.7, 4), (2500, 0.1, 5), (33, 0.3, 6)])
def test_nms_lazy_equals_full_mask_equals_oracle(n, thresh, seed):
    """the default (kept rows only) and the mask + sweep formulation give the identical keep lists"""
    boxes, scores = synth.nms_frames(3, n, seed=500 + seed)
    # one frame of near-duplicates: exercises the literal (atan2) path for > 8 vertices and angular ties
    r = np.random.default_rng(seed)
    boxes[2] = boxes[2][r.integers(0, max(n // 8, 1), n)] + r.normal(0, 2e-3, (n, 7)).astype(np.float32)
    tb, ts = cu(boxes), cu(scores)
    k_lazy, n_lazy = U.nms_gpu_batched(tb, ts, thresh)
    k_full, n_full = U.nms_gpu_batched(tb, ts, thresh, full_mask=True)
    assert torch.equal(n_lazy, n_full) and torch.equal(k_lazy, k_full)
    # small batches spread every problem over a thread-block cluster (2-8 SMs, DSMEM exchange); one CTA per problem must agree
    k_one, n_one = U._nms_batched('lg_nms_rotated_batched', tb, ts, thresh, None, flags=_lib.LG_FLAG_NMS_NO_CLUSTER)
    assert torch.equal(n_lazy, n_one) and torch.equal(k_lazy, k_one)
    for f in range(3):
        order = ts[f].sort(0, descending=True)[1].cpu().numpy()
        want = O.nms(boxes[f], scores[f], thresh, flavor=O.FLAVOR_CUDA, order=order)
        got = k_lazy[f, : int(n_lazy[f])].cpu().numpy()
        inv = np.empty(n, np.int64)
        inv[order] = np.arange(n)
        if f < 2 or not R.available():
            assert np.array_equal(got, want), explain_nms_mismatch(boxes[f][order], inv[got], inv[want], thresh, False)
        else:
            # near-duplicates sit on vertex-order ties, where the oracle's glibc atan2f is not the authority: the reference
            # kernel itself (libdevice atan2f), run live on this GPU, is (iou3d_nms.cpp:90-136)
            bs = tb[f][torch.from_numpy(order).to(dev())].contiguous()
            keep = torch.LongTensor(n)
            nk = R.iou3d_nms_cuda().nms_gpu(bs, keep, thresh)
            want_live = order[keep[:nk].numpy()]
            assert np.array_equal(got, want_live), explain_nms_mismatch(boxes[f][order], inv[got], inv[want_live], thresh, False)


def test_cfg5_multihead_full_size_small_problem_variant():
    """BASELINE configs[4] at full size: NuScenes CBGS multi-head NMS, 256 frames x 10 classes = 2560 problems of 1000 boxes,
    thresh 0.2 (cbgs_second_multihead.yaml:196-206, model_nms_utils.py:28-65).  2560 >= 2 x SMs problems of <= 1536 boxes
    launch nms_lazy_kernel<., false, 256> (two 256-thread CTAs per SM) behind torch's segmented sort -- the variant
    bench.py --workload nms_cfg5 times; sub-batches of 160 problems take the 512-thread / lg_select_topk variant."""
    b, s = synth.cfg5(256, 10, 1000, seed=synth.SEEDS["cfg5"])
    boxes, scores = b.reshape(-1, 1000, 7), s.reshape(-1, 1000)
    P = boxes.shape[0]
    assert P == 2560
    tb, ts = cu(boxes), cu(scores)
    keep, num = U.nms_gpu_batched(tb, ts, 0.2)
    assert keep.shape == (P, 1000) and num.shape == (P,)
    num_h, keep_h = num.cpu().numpy(), keep.cpu().numpy()
    # (1) the oracle on a stride of problems
    for p in range(0, P, 97):
        order = ts[p].sort(dim=0, descending=True, stable=True)[1].cpu().numpy()
        want = O.nms(boxes[p], scores[p], 0.2, flavor=O.FLAVOR_CUDA, order=order)
        got = keep_h[p, : num_h[p]]
        inv = np.empty(1000, np.int64)
        inv[order] = np.arange(1000)
        assert np.array_equal(got, want), (p, explain_nms_mismatch(boxes[p][order], inv[got], inv[want], 0.2, False))
        assert (keep_h[p, num_h[p]:] == -1).all()
    # (2) the 512-thread variant (a batch below 2 x SMs problems) and the full-mask formulation on sub-batches
    for p0 in (0, 1200, 2400):
        k1, n1 = U.nms_gpu_batched(tb[p0:p0 + 160], ts[p0:p0 + 160], 0.2)
        assert torch.equal(n1, num[p0:p0 + 160]) and torch.equal(k1, keep[p0:p0 + 160])
    k2, n2 = U.nms_gpu_batched(tb[:320], ts[:320], 0.2, full_mask=True)
    assert torch.equal(n2, num[:320]) and torch.equal(k2, keep[:320])
    # (3) ragged counts through the same variant
    counts = torch.from_numpy(np.random.default_rng(1).integers(0, 1001, P).astype(np.int32))
    kc, nc = U.nms_gpu_batched(tb, ts, 0.2, counts)
    for p in range(0, P, 211):
        c = int(counts[p])
        want = U.nms_gpu(tb[p, :c], ts[p, :c], 0.2)[0]
        assert int(nc[p]) == want.numel() and torch.equal(kc[p, : want.numel()], want)


@pytest.mark.parametrize("thresh", [0.01, 0.2, 0.7])
def test_nms_post_maxsize_is_the_truncated_keep_list(thresh):
    """max_keep = NMS_POST_MAXSIZE (model_nms_utils.py:20 `selected[:NMS_POST_MAXSIZE]`): the first max_keep entries of the full
    keep list, for the lazy kernel (which stops early; independent candidates are kept out of score order, so "early" has to
    wait until that many are kept below the cursor), the mask + sweep formulation and the axis-aligned NMS; any row pitch"""
    boxes, scores = synth.nms_frames(5, 1500, seed=31)
    counts = torch.tensor([1500, 700, 0, 33, 1500], dtype=torch.int32)
    tb, ts = cu(boxes), cu(scores)
    full_k, full_n = U.nms_gpu_batched(tb, ts, thresh, counts)
    norm_k, norm_n = U.nms_normal_gpu_batched(tb, ts, thresh, counts)
    for mk in (1, 7, 40, 500, 5000):
        K = min(mk, 1500)
        for fn, wk, wn, kw in ((U.nms_gpu_batched, full_k, full_n, {}), (U.nms_gpu_batched, full_k, full_n, {"full_mask": True}),
                               (U.nms_normal_gpu_batched, norm_k, norm_n, {})):
            k, n = fn(tb, ts, thresh, counts, max_keep=mk, **kw)
            assert k.shape == (5, K)
            assert torch.equal(n, torch.clamp(wn, max=K)) and torch.equal(k, wk[:, :K]), (mk, kw)
    # caller-owned keep with a wider row pitch (the packed send buffer of the sharded gather): only the K columns are written
    send = torch.full((5, 1 + 40), -7, dtype=torch.int64, device=dev())
    k, n = U.nms_gpu_batched(tb, ts, thresh, counts, max_keep=40, keep_out=send[:, 1:])
    assert k.data_ptr() == send[:, 1:].data_ptr() and torch.equal(send[:, 1:], full_k[:, :40]) and bool((send[:, 0] == -7).all())
    assert float(full_n.float().mean()) > 5


@pytest.mark.parametrize("n", [20000, 52000])
def test_nms_large_problems(n):
    """20,000 boxes: lazy kernel with the cull quads beyond its 4096-entry smem cache; 52,000: the alive / suppression
    rows no longer fit in shared memory and the mask + sweep formulation takes over -- same keep list either way"""
    boxes, scores = synth.nms_frames(1, n, seed=n, k_range=(200, 300))
    tb, ts = cu(boxes[0]), cu(scores[0])
    keep = U.nms_gpu(tb, ts, 0.05)[0].cpu().numpy()
    order = ts.sort(0, descending=True)[1].cpu().numpy()
    want = O.nms(boxes[0], scores[0], 0.05, flavor=O.FLAVOR_CUDA, order=order)
    assert np.array_equal(keep, want), (len(keep), len(want))


def test_nms_vs_golden_reference_cuda(ggpu):
    boxes, scores = ggpu["nms_boxes"], ggpu["nms_scores"]
    for f in range(boxes.shape[0]):
        tb, ts = cu(boxes[f]), cu(scores[f])
        assert np.array_equal(ts.sort(0, descending=True)[1].cpu().numpy(), ggpu[f"nms_order_{f}"])
        for thr in (0.01, 0.1, 0.7):
            assert np.array_equal(U.nms_gpu(tb, ts, thr)[0].cpu().numpy(), ggpu[f"nms_keep_{f}_{thr}_0"]), (f, thr)
            assert np.array_equal(U.nms_normal_gpu(tb, ts, thr)[0].cpu().numpy(), ggpu[f"nms_keep_{f}_{thr}_1"]), (f, thr)


def test_nms_from_pinned_host_inputs_equals_the_device_call():
    boxes, scores = synth.nms_frames(6, 900, seed=61)
    hb, hs = torch.from_numpy(boxes).pin_memory(), torch.from_numpy(scores).pin_memory()
    counts = torch.tensor([900, 1, 0, 64, 65, 333], dtype=torch.int32)
    for c in (None, counts):
        for mk in (None, 50):
            k0, n0 = U.nms_gpu_batched(cu(boxes), cu(scores), 0.1, c, max_keep=mk)
            for _ in range(3):  # the side stream and its buffers are reused across calls
                k1, n1 = U.nms_gpu_batched_from_host(hb, hs, 0.1, c, max_keep=mk)
                assert torch.equal(k0, k1) and torch.equal(n0, n1)


def test_host_nms_pipeline_equals_the_device_call_batch_by_batch():
    """HostNmsPipeline: batches in flight share nothing -- every ticket returns the keep lists of ITS batch, in any collection
    order the depth allows, and a slot is not reused before its result was collected"""
    batches = [synth.nms_frames(5, 700, seed=70 + k) for k in range(5)]
    host = [(torch.from_numpy(b).pin_memory(), torch.from_numpy(s).pin_memory()) for b, s in batches]
    want = [U.nms_gpu_batched(cu(b), cu(s), 0.1, max_keep=40) for b, s in batches]
    for depth in (1, 2, 3):
        pipe = U.HostNmsPipeline(5, 700, 0.1, max_keep=40, depth=depth)
        tickets = []
        for k, (hb, hs) in enumerate(host):
            if len(tickets) == depth:  # collect the oldest before its slot is needed again
                t = tickets.pop(0)
                hk, hn = pipe.result(t)
                assert torch.equal(hk, want[t][0].cpu()) and torch.equal(hn, want[t][1].cpu()), (depth, t)
            tickets.append(pipe.submit(hb, hs))
        for t in tickets:
            hk, hn = pipe.result(t)
            assert torch.equal(hk, want[t][0].cpu()) and torch.equal(hn, want[t][1].cpu()), (depth, t)
    pipe = U.HostNmsPipeline(5, 700, 0.1, max_keep=40, depth=1)
    pipe.submit(*host[0])
    with pytest.raises(RuntimeError):
        pipe.submit(*host[1])  # depth 1: the first result has not been collected
    with pytest.raises(RuntimeError):
        pipe.result(7)


def test_nms_of_more_than_65536_boxes():
    """no pre_maxsize, raw candidates: the reference takes any N (it allocates the N x N/64 mask); so does the mask + sweep
    formulation here (rotated and axis-aligned), checked against the oracle's NMS"""
    boxes, scores = synth.nms_frames(1, 70001, seed=77)
    tb, ts = cu(boxes[0]), cu(scores[0])
    order = ts.sort(dim=0, descending=True, stable=True)[1].cpu().numpy()
    for fn, normal, thr in ((U.nms_gpu, False, 0.1), (U.nms_normal_gpu, True, 0.3)):
        got = fn(tb, ts, thr)[0].cpu().numpy()
        want = O.nms(boxes[0], scores[0], thr, normal=normal, order=order)
        assert np.array_equal(got, want), (normal, len(got), len(want))


def test_nms_threshold_edge_values():
    """thresh < 0: the exact-zero IoU of disjoint boxes exceeds it too (kernel.cu:304), so only the best box survives -- the
    kernels' exact-zero cull must not change that; thresh = NaN: nothing is suppressed; thresh >= 1: only IoUs above 1"""
    boxes, scores = synth.nms_frames(3, 300, seed=41)
    tb, ts = cu(boxes), cu(scores)
    counts = torch.tensor([300, 0, 17], dtype=torch.int32)
    for thr in (-0.5, -1e-9):
        for fn, normal in ((U.nms_gpu_batched, False), (U.nms_normal_gpu_batched, True)):
            keep, num = fn(tb, ts, thr, counts)
            for p in range(3):
                c = int(counts[p])
                order = ts[p, :c].sort(0, descending=True)[1].cpu().numpy()
                want = O.nms(boxes[p, :c], scores[p, :c], thr, normal=normal, order=order) if c else np.zeros(0, np.int64)
                assert len(want) == min(c, 1) and np.array_equal(keep[p, : int(num[p])].cpu().numpy(), want), (thr, normal, p)
                assert bool((keep[p, int(num[p]):] == -1).all())
        k1 = U.nms_gpu(tb[0], ts[0], thr)[0]
        assert k1.numel() == 1 and int(k1[0]) == int(ts[0].argmax())
    keep, num = U.nms_gpu_batched(tb, ts, float("nan"))
    assert num.tolist() == [300, 300, 300]
    keep, num = U.nms_gpu_batched(tb, ts, 1.5)
    assert num.tolist() == [300, 300, 300]


def test_nms_pre_maxsize_empty_and_duplicates():
    boxes, scores = synth.nms_frames(1, 500, seed=9)
    tb, ts = cu(boxes[0]), cu(scores[0])
    order = ts.sort(0, descending=True)[1].cpu().numpy()
    got = U.nms_gpu(tb, ts, 0.1, pre_maxsize=200)[0].cpu().numpy()
    assert np.array_equal(got, O.nms(boxes[0], scores[0], 0.1, pre_maxsize=200, order=order))
    k, _ = U.nms_gpu(tb[:0], ts[:0], 0.1)
    assert k.shape == (0,) and k.dtype == torch.int64
    one = np.array([[1, 2, 0, 4, 2, 1, 0.3]], np.float32)
    dup = cu(np.repeat(one, 130, 0))
    s = cu(np.linspace(1, 0.1, 130).astype(np.float32))
    assert U.nms_gpu(dup, s, 0.5)[0].cpu().tolist() == [0]
    assert U.nms_gpu(dup, s, 1.0)[0].cpu().tolist() == list(range(130))  # strict >
    assert U.nms_normal_gpu(dup, s, 0.5)[0].cpu().tolist() == [0]


def test_nms_batched_with_ragged_counts():
    boxes, scores = synth.nms_frames(6, 700, seed=12)
    counts = torch.tensor([700, 1, 0, 64, 65, 333], dtype=torch.int32)
    tb, ts = cu(boxes), cu(scores)
    for fn, single, normal in ((U.nms_gpu_batched, U.nms_gpu, False), (U.nms_normal_gpu_batched, U.nms_normal_gpu, True)):
        keep, num = fn(tb, ts, 0.1, counts)
        assert keep.shape == (6, 700) and num.dtype == torch.int32
        for p in range(6):
            c = int(counts[p])
            want = single(tb[p, :c], ts[p, :c], 0.1)[0]
            assert int(num[p]) == want.numel()
            assert torch.equal(keep[p, : want.numel()], want)
            assert bool((keep[p, want.numel():] == -1).all())
            order = ts[p, :c].sort(0, descending=True)[1].cpu().numpy()
            assert np.array_equal(want.cpu().numpy(), O.nms(boxes[p, :c], scores[p, :c], 0.1, normal=normal, order=order))


# ------------------------------------------------------------------------------------------ points
def test_points_in_boxes_vs_oracle_and_golden(ggpu):
    pts, rois = synth.cfg3(n_frames=3, n_points=16384, n_rois=100, seed=5)
    got = PU.points_in_boxes_gpu(cu(pts), cu(rois))
    assert got.dtype == torch.int32 and got.shape == (3, 16384)
    want = O.points_in_boxes_idx(pts, rois, O.FLAVOR_CUDA)
    assert np.array_equal(got.cpu().numpy(), want)
    assert 0.2 < (want >= 0).mean() < 0.5
    g = PU.points_in_boxes_gpu(cu(ggpu["pib_pts"]), cu(ggpu["pib_boxes"])).cpu().numpy()
    assert np.array_equal(g, ggpu["pib_idx"])


@pytest.mark.parametrize("T,M,spread", [(200, 12345, 0.08), (254, 5000, 1.0), (255, 3001, 1.0), (300, 2048, 1.0), (3, 70001, 1.0)])
def test_points_grid_paths_vs_oracle(T, M, spread):
    """compact cell lists incl. overflowing cells (tight cluster), the 254-box boundary, the test-every-box path,
    point counts that are not multiples of 4 (fallback tile loads) and frames split over several CTAs"""
    r = np.random.default_rng(T + M)
    B = 3
    boxes = np.stack([synth.gt_boxes(T, int(r.integers(1 << 30))) for _ in range(B)])
    c = boxes[:, :, 0:2].mean(1, keepdims=True)
    boxes[:, :, 0:2] = c + (boxes[:, :, 0:2] - c) * spread
    pts = np.empty((B, M, 3), np.float32)
    for b in range(B):
        k = r.integers(0, T, M)
        loc = r.uniform(-0.6, 0.6, (M, 3)) * boxes[b, k, 3:6]
        co, si = np.cos(boxes[b, k, 6]), np.sin(boxes[b, k, 6])
        pts[b, :, 0] = boxes[b, k, 0] + loc[:, 0] * co - loc[:, 1] * si
        pts[b, :, 1] = boxes[b, k, 1] + loc[:, 0] * si + loc[:, 1] * co
        pts[b, :, 2] = boxes[b, k, 2] + loc[:, 2]
    got = PU.points_in_boxes_gpu(cu(pts), cu(boxes)).cpu().numpy()
    want = O.points_in_boxes_idx(pts, boxes, O.FLAVOR_CUDA)
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} differ"


def test_points_edge_cases():
    # ragged sizes, padded all-zero boxes (dataset.py:172-177 pads GT with zeros; they are NOT skipped)
    pts = np.zeros((2, 1001, 3), np.float32)
    pts[:, 1:] = np.random.default_rng(3).uniform(-5, 5, (2, 1000, 3))
    boxes = np.zeros((2, 7, 7), np.float32)
    boxes[:, :3] = synth.gt_boxes(6, 2, x_range=(-4, 4), y_range=(-4, 4), z=0.0).reshape(2, 3, 7)
    got = PU.points_in_boxes_gpu(cu(pts), cu(boxes)).cpu().numpy()
    want = O.points_in_boxes_idx(pts, boxes, O.FLAVOR_CUDA)
    assert np.array_equal(got, want)
    assert got[0, 0] >= 0  # the origin matches a zero-size padded box (or an earlier real one)
    assert PU.points_in_boxes_gpu(cu(pts[:, :0]), cu(boxes)).shape == (2, 0)
    none = PU.points_in_boxes_gpu(cu(pts), cu(boxes[:, :0])).cpu().numpy()
    assert (none == -1).all()
    # x/y open with 1e-5 margin evaluated in double, z closed
    box = np.array([[[0, 0, 0, 4, 2, 1.5, 0.0], [0, 0, 0, 4, 2, 1.5, 0.0]]], np.float32)
    p = np.array([[[0, 0, 0], [2.000005, 0, 0], [2.00002, 0, 0], [0, 0, 0.75], [0, 0, 0.7500001], [2.005, 0, 0]]], np.float32)
    assert PU.points_in_boxes_gpu(cu(p), cu(box)).cpu().tolist() == [[0, 0, -1, 0, -1, -1]]


def test_sizes_beyond_the_kernels_grid_and_shared_memory_limits():
    """the reference entry points take any batch / box count; the library's limits (65,535 frames or problems per launch, 2048
    boxes per frame in shared memory) are handled by slicing, not by raising"""
    r = np.random.default_rng(12)
    # 5000 boxes per frame: chunks of 2048, the lowest-index hit wins across chunks
    boxes = synth.gt_boxes(5000, 3)[None]
    k = r.integers(0, 5000, 6000)
    pts = (boxes[0, k, 0:3] + r.normal(0, 0.4, (6000, 3))).astype(np.float32)[None]
    got = PU.points_in_boxes_gpu(cu(pts), cu(boxes)).cpu().numpy()
    assert np.array_equal(got, O.points_in_boxes_idx(pts, boxes, O.FLAVOR_CUDA)) and (got >= 2048).any() and (got >= 4096).any()
    # 70,000 tiny frames
    B = 70000
    fb = np.tile(synth.gt_boxes(3, 5)[None], (B, 1, 1))
    fb[:, :, 0] += r.normal(0, 0.5, (B, 3)).astype(np.float32)
    fp = (fb[:, r.integers(0, 3, 8), 0:3] + r.normal(0, 0.8, (B, 8, 3))).astype(np.float32)
    got = PU.points_in_boxes_gpu(cu(fp), cu(fb)).cpu().numpy()
    rows = np.concatenate([np.arange(0, B, 997), np.arange(65500, 65600), [B - 1]])
    assert np.array_equal(got[rows], O.points_in_boxes_idx(fp[rows], fb[rows], O.FLAVOR_CUDA))
    # 70,000 NMS problems of 12 boxes, rotated and axis-aligned
    nb, ns = synth.nms_frames(64, 12, seed=4)
    nb, ns = np.tile(nb, (1100, 1, 1))[:B], np.tile(ns, (1100, 1))[:B]
    for fn in (U.nms_gpu_batched, U.nms_normal_gpu_batched):
        keep, num = fn(cu(nb), cu(ns), 0.1)
        k64, n64 = fn(cu(nb[:64]), cu(ns[:64]), 0.1)
        assert torch.equal(keep[:1093 * 64].view(1093, 64, 12), k64.unsqueeze(0).expand(1093, 64, 12))  # the batch repeats every 64 problems
        assert torch.equal(num[:1093 * 64].view(1093, 64), n64.unsqueeze(0).expand(1093, 64))
        assert torch.equal(keep[65536:65600], k64) and torch.equal(num[65536:65600], n64)  # 65536 = 1024 x 64: the tiling restarts there


def test_cpu_named_functions_run_on_the_gpu_with_cpu_semantics():
    g = np.load(os.path.join(HERE, "golden", "golden_cpu.npz"))
    for f in range(2):
        m = PU.points_in_boxes_cpu(g["pib_pts"][f], g["pib_boxes"][f])
        assert isinstance(m, np.ndarray) and m.dtype == np.int32
        assert np.array_equal(m, g["pib_ref_mask"][f])  # bit-exact vs the reference's own CPU output
    t = PU.points_in_boxes_cpu(torch.from_numpy(g["pib_pts"][0]), torch.from_numpy(g["pib_boxes"][0]))
    assert isinstance(t, torch.Tensor) and not t.is_cuda
    for name in IOU_SETS:  # all nine sets of the reference's compiled boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252), 1e-5 absolute
        a, b, ref = g[f"iou_{name}_a"], g[f"iou_{name}_b"], g[f"iou_{name}_ref"]
        got = U.boxes_bev_iou_cpu(a, b)
        assert isinstance(got, np.ndarray) and got.shape == ref.shape
        nbad = assert_iou_close(got, ref, f"boxes_bev_iou_cpu {name}")
        # bit for bit: the strict flavor restates the CPU build's un-contracted FP32 AND its glibc sinf / cosf / atan2f (lg_trig.cuh),
        # so even the hand-made tie sets (kat*: identical, nested and edge-sharing boxes) come out identical
        assert nbad == 0, f"{name}: {nbad} of {ref.size} entries not bit-identical to the reference CPU build"


# ------------------------------------------------------------------------------------------ tier A, live
@pytest.mark.skipif(not R.available(), reason="oracle/_ref (compiled reference) did not travel with this snapshot")
def test_live_against_reference_cuda_kernels():
    ref, roi = R.iou3d_nms_cuda(), R.roiaware_pool3d_cuda()
    tot = bad = 0
    for seed, centre, pri in [(21, (35, 17.5), synth.KITTI_PRIORS), (22, (70, 35), synth.KITTI_PRIORS[1:2]),
                              (23, (150, 75), synth.KITTI_PRIORS), (24, (-60, 40), synth.WAYMO_PRIORS)]:
        a, b = synth.clustered_pairs(1000, 1000, seed, centre, pri)
        ta, tb = cu(a), cu(b)
        want = torch.zeros((1000, 1000), device=dev())
        ref.boxes_iou_bev_gpu(ta, tb, want)
        got = U.boxes_iou_bev(ta, tb)
        bad += assert_iou_close(got.cpu().numpy(), want.cpu().numpy(), "tier A iou_bev")
        tot += 1000 * 1000
    assert bad <= tot // 2000, f"{bad} of {tot} pairs not bit-identical to the reference CUDA kernel"
    boxes, scores = synth.cfg2(n_frames=3, n_boxes=4096)
    for f in range(3):
        tb, ts = cu(boxes[f]), cu(scores[f])
        order = ts.sort(0, descending=True)[1]
        b = tb[order].contiguous()
        for thr, normal in ((0.01, False), (0.5, False), (0.01, True)):
            keep = torch.LongTensor(b.size(0))
            n = (ref.nms_normal_gpu if normal else ref.nms_gpu)(b, keep, thr)
            want = order[keep[:n].to(dev())]
            got = (U.nms_normal_gpu if normal else U.nms_gpu)(tb, ts, thr)[0]
            assert torch.equal(got, want), (f, thr, normal)
    pts, rois = synth.cfg3(n_frames=4, seed=9)
    want = torch.full((4, 16384), -1, dtype=torch.int32, device=dev())
    roi.points_in_boxes_gpu(cu(rois), cu(pts), want)
    assert torch.equal(PU.points_in_boxes_gpu(cu(pts), cu(rois)), want)


def _adversarial_boxes(r, n):
    """extreme aspect ratios, tiny and huge boxes, far from the origin, headings at exact multiples of pi/2, exact duplicates,
    exactly touching and nested boxes, near-duplicates a few ulp / mm apart"""
    b = synth.gt_boxes(n, int(r.integers(1 << 30)), x_range=(-5, 5), y_range=(-5, 5))
    k = n // 10
    b[0 * k:1 * k, 3] = r.uniform(8, 30, k); b[0 * k:1 * k, 4] = r.uniform(0.05, 0.3, k)          # needles
    b[1 * k:2 * k, 3:5] = r.uniform(0.02, 0.08, (k, 2))                                            # tiny (margin dominates)
    b[2 * k:3 * k, 3:5] = r.uniform(40, 120, (k, 2))                                               # huge
    b[3 * k:4 * k, 0:2] += np.array([1000.0, -2000.0], np.float32)                                # far away (coarse ulp)
    b[4 * k:5 * k, 6] = r.integers(-4, 5, k) * np.float32(np.pi / 2)                               # axis aligned, on a 0.5 m lattice
    b[4 * k:5 * k, 0:2] = np.round(b[4 * k:5 * k, 0:2] * 2) / 2
    b[4 * k:5 * k, 3:5] = np.round(b[4 * k:5 * k, 3:5] * 2 + 1) / 2
    b[5 * k:6 * k] = b[4 * k:5 * k]                                                                # exact duplicates of those
    b[6 * k:7 * k] = b[0:k]
    b[6 * k:7 * k, 0:2] += r.normal(0, 1e-3, (k, 2)).astype(np.float32)                           # near-duplicates (mm)
    b[7 * k:8 * k] = b[k:2 * k]
    b[7 * k:8 * k, 6] += r.normal(0, 1e-4, k).astype(np.float32)                                  # near-duplicates (angle)
    b[8 * k:9 * k, 6] = 0.0
    b[8 * k:9 * k, 0] = b[8 * k, 0] + np.arange(k) * b[8 * k, 3]                                   # a row of boxes touching edge to edge
    b[8 * k:9 * k, 1] = b[8 * k, 1]
    b[8 * k:9 * k, 3:5] = b[8 * k, 3:5]
    return b.astype(np.float32)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (compiled reference) did not travel with this snapshot")
def test_adversarial_boxes_live_against_reference_cuda_kernels():
    """the reference CUDA kernels, run live on this GPU, are the authority: every entry bit-identical or a vertex-order tie"""
    ref, roi = R.iou3d_nms_cuda(), R.roiaware_pool3d_cuda()
    r = np.random.default_rng(2024)
    a, b = _adversarial_boxes(r, 600), _adversarial_boxes(r, 500)
    b[:200] = a[:200]  # shared boxes: identical, touching and nested pairs across the two sets
    ta, tb = cu(a), cu(b)
    want = torch.zeros((600, 500), device=dev())
    ref.boxes_iou_bev_gpu(ta, tb, want)
    got = U.boxes_iou_bev(ta, tb).cpu().numpy()
    want = want.cpu().numpy()
    assert np.array_equal(got == 0, want == 0), "zero pattern differs"
    diff = bits(got) != bits(want)
    # the only licensed difference: polygons with two vertices at (numerically) the same polar angle, whose order hangs on the
    # last bit of atan2f -- libdevice's in the reference kernel, and also in ours (overlap_area_slow), so none is expected
    assert int(diff.sum()) <= 3 and np.abs(got - want).max() <= 2e-3, (int(diff.sum()), float(np.abs(got - want).max()))
    wov = torch.zeros((600, 500), device=dev())
    ref.boxes_overlap_bev_gpu(ta, tb, wov)
    gov = U.boxes_overlap_bev(ta, tb)
    assert int((gov != wov).sum()) <= 3
    # NMS on the adversarial set, rotated and axis aligned
    sc = cu(r.permutation(np.linspace(0.1, 1.0, 600)).astype(np.float32))
    order = sc.sort(0, descending=True)[1]
    bs = ta[order].contiguous()
    for thr, normal in ((0.05, False), (0.5, False), (0.9, False), (0.3, True)):
        keep = torch.LongTensor(600)
        nk = (ref.nms_normal_gpu if normal else ref.nms_gpu)(bs, keep, thr)
        wantk = order[keep[:nk].to(dev())]
        gotk = (U.nms_normal_gpu if normal else U.nms_gpu)(ta, sc, thr)[0]
        assert torch.equal(gotk, wantk), (thr, normal, len(gotk), nk)
    # points on and around the faces of adversarial boxes
    boxes = a[:120][None]
    pts = np.concatenate([r.uniform(-8, 8, (20000, 3)), boxes[0, r.integers(0, 120, 5000), 0:3] + r.normal(0, 0.5, (5000, 3))]).astype(np.float32)[None]
    wantp = torch.full((1, pts.shape[1]), -1, dtype=torch.int32, device=dev())
    roi.points_in_boxes_gpu(cu(boxes), cu(pts), wantp)
    assert torch.equal(PU.points_in_boxes_gpu(cu(pts), cu(boxes)), wantp)


# ------------------------------------------------------------------------------------------ post-processing front end (8f-1)
def _ref_class_agnostic_nms(box_scores, box_preds, cfg, score_thresh):
    """the reference's model_nms_utils.class_agnostic_nms, restated around the ORACLE's NMS (numpy)"""
    idx = np.nonzero(box_scores >= score_thresh)[0] if score_thresh is not None else np.arange(len(box_scores))
    s, b = box_scores[idx], box_preds[idx]
    if len(s) == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(-s, kind="stable")[: min(cfg["NMS_PRE_MAXSIZE"], len(s))]
    keep = O.nms(b[order][:, :7], s[order], cfg["NMS_THRESH"], normal=cfg["NMS_TYPE"] == "nms_normal_gpu", flavor=O.FLAVOR_CUDA)
    return idx[order[keep[: cfg["NMS_POST_MAXSIZE"]]]]


@pytest.mark.parametrize("nms_type", ["nms_gpu", "nms_normal_gpu"])
def test_post_processing_front_end_batched_and_per_frame(nms_type):
    from lidardetection_b200 import model_nms_utils as MU

    cfg = {"NMS_TYPE": nms_type, "NMS_THRESH": 0.1, "NMS_PRE_MAXSIZE": 512, "NMS_POST_MAXSIZE": 40, "MULTI_CLASSES_NMS": False}
    B, N = 5, 3000
    boxes, scores = synth.nms_frames(B, N, seed=77)
    boxes9 = np.concatenate([boxes, np.zeros((B, N, 2), np.float32)], 2)  # 7 + C columns, as the detectors pass them
    scores[4] *= 0.05  # a frame where nothing passes the score threshold
    tb, ts = cu(boxes9), cu(scores)
    sel, num, sc = MU.class_agnostic_nms_batched(ts, tb, cfg, score_thresh=0.3)
    assert sel.shape == (B, 40) and num.dtype == torch.int32
    for f in range(B):
        want = _ref_class_agnostic_nms(scores[f], boxes9[f], cfg, 0.3)
        got = sel[f, : int(num[f])].cpu().numpy()
        assert np.array_equal(got, want), f
        assert bool((sel[f, int(num[f]):] == -1).all())
        one, one_scores = MU.class_agnostic_nms(ts[f], tb[f], cfg, score_thresh=0.3)  # the reference's per-frame entry point
        assert np.array_equal(one.cpu().numpy(), want)
        assert torch.equal(one_scores, ts[f][one]) and torch.equal(sc[f, : int(num[f])], ts[f][one])
    assert int(num[4]) == 0


def test_multi_classes_nms_batched_matches_per_class_oracle():
    from lidardetection_b200 import model_nms_utils as MU

    cfg = {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.2, "NMS_PRE_MAXSIZE": 300, "NMS_POST_MAXSIZE": 25}
    B, N, C = 2, 900, 3
    boxes, _ = synth.nms_frames(B, N, seed=5)
    r = np.random.default_rng(9)
    cls = r.permuted(np.tile(np.linspace(0.01, 0.99, N * C, dtype=np.float32), (B, 1)), axis=1).reshape(B, N, C)
    sel, num, _ = MU.multi_classes_nms_batched(cu(cls), cu(boxes), cfg, score_thresh=0.2)
    for b in range(B):
        for k in range(C):
            want = _ref_class_agnostic_nms(cls[b, :, k], boxes[b], cfg, 0.2)
            assert np.array_equal(sel[b, k, : int(num[b, k])].cpu().numpy(), want), (b, k)
    ps, pl, pb = MU.multi_classes_nms(cu(cls[0]), cu(boxes[0]), cfg, score_thresh=0.2)  # the reference's per-frame signature
    want_idx = np.concatenate([_ref_class_agnostic_nms(cls[0, :, k], boxes[0], cfg, 0.2) for k in range(C)])
    assert np.array_equal(pb.cpu().numpy(), boxes[0][want_idx]) and len(ps) == len(pl) == len(want_idx)


# ------------------------------------------------------------------------------------------ fused row / column maxima (8f-4)
@pytest.mark.parametrize("kind,fn", [("iou3d", "boxes_iou3d_gpu"), ("iou_bev", "boxes_iou_bev"), ("overlap_bev", "boxes_overlap_bev")])
@pytest.mark.parametrize("n,m,sparse", [(300, 257, False), (1000, 20, True), (2500, 3100, True), (64, 1, False)])
def test_iou_row_col_max_equals_torch_max_of_the_matrix(kind, fn, n, m, sparse):
    if sparse:
        a, b = synth.cfg4(max(n, m), seed=n + m)
        a, b = a[:n], b[:m]
    else:
        a, b = synth.clustered_pairs(n, m, seed=n)
    ta, tb = cu(a), cu(b)
    mat = getattr(U, fn)(ta, tb)
    rmax, rarg, cmax, carg = U.boxes_iou_max(ta, tb, kind=kind, rows=True, cols=True)
    wr, wc = mat.max(1), mat.max(0)
    assert torch.equal(rmax, wr.values) and torch.equal(cmax, wc.values)  # bit-identical values
    # argmax = lowest index attaining the maximum (torch documents the same convention)
    first_r = (mat == wr.values.unsqueeze(1)).int().argmax(1)
    first_c = (mat == wc.values.unsqueeze(0)).int().argmax(0)
    assert torch.equal(rarg, first_r) and torch.equal(carg, first_c)
    only_rows = U.boxes_iou_max(ta, tb, kind=kind)
    assert len(only_rows) == 2 and torch.equal(only_rows[0], rmax) and torch.equal(only_rows[1], rarg)


def test_two_phase_sweep_and_its_overflow_path_equal_the_one_kernel_sweep():
    """From 2^26 pairs on the matrix is swept in two phases (zeros + survivor list, then the polygon path on the list).  The same
    matrix must come out (a) of the one-kernel sweep (LG_FLAG_IOU_ONE_KERNEL) and (b) when the list is far too small
    (LG_FLAG_IOU_SMALL_LIST: 1024 entries), where the flagged strips are redone by the one-kernel sweep; likewise the reductions."""
    a, b = synth.cfg4(n=9000, seed=11)  # 8.1e7 pairs, some 1e5 survivors
    ta, tb = cu(a), cu(b)
    ONE, SMALL = _lib.LG_FLAG_IOU_ONE_KERNEL, _lib.LG_FLAG_IOU_SMALL_LIST
    for fn in ("lg_boxes_iou3d", "lg_boxes_iou_bev", "lg_boxes_overlap_bev"):
        two = U._iou_call(fn, ta, tb)
        assert int((two > 0).sum()) > 4096  # more overlapping pairs than the small list holds
        one = U._iou_call(fn, ta, tb, flags=ONE)
        assert torch.equal(two, one), fn
        small = U._iou_call(fn, ta, tb, flags=SMALL)
        assert torch.equal(small, one), fn
    # ragged shapes: partial row strips, partial column tiles, a pitch that rules out the 16-byte stores
    n, m = 8999, 8195
    wide = torch.full((n, m + 3), -7.0, dtype=torch.float32, device=ta.device)
    got = U._iou_call("lg_boxes_iou3d", ta[:n], tb[:m], out=wide[:, :m])
    want = U._iou_call("lg_boxes_iou3d", ta[:n], tb[:m], flags=ONE)
    assert torch.equal(got, want) and bool((wide[:, m:] == -7.0).all())
    # the reductions over the same pairs
    w = U._iou_call("lg_boxes_iou3d", ta, tb, flags=ONE)
    for flags in (_lib.LG_FLAG_NONE, SMALL, ONE):
        rmax, rarg, cmax, carg = U.boxes_iou_max(ta, tb, kind="iou3d", rows=True, cols=True, flags=flags)
        assert torch.equal(rmax, w.max(1).values) and torch.equal(cmax, w.max(0).values)
        assert torch.equal(rarg, (w == w.max(1).values.unsqueeze(1)).int().argmax(1))
        assert torch.equal(carg, (w == w.max(0).values.unsqueeze(0)).int().argmax(0))
    # a dense matrix above the threshold: the density probe routes it to the dense build, the two-phase kernels stand down
    da, db = synth.clustered_pairs(8200, 8200, seed=5)
    tda, tdb = cu(da), cu(db)
    dense = U._iou_call("lg_boxes_iou_bev", tda, tdb)
    assert torch.equal(dense, U._iou_call("lg_boxes_iou_bev", tda, tdb, flags=ONE))
    rmax, rarg = U.boxes_iou_max(tda, tdb, kind="iou_bev")  # no probe on this path: the list overflows, every strip falls back
    assert torch.equal(rmax, dense.max(1).values)


def test_iou_row_max_at_a_size_whose_matrix_is_never_built():
    """60,000 x 60,000 (a 14.4 GB matrix): reduce it, then check sampled rows against the matrix entry point"""
    a, b = synth.cfg4(60000, seed=3)
    ta, tb = cu(a), cu(b)
    rmax, rarg = U.boxes_iou_max(ta, tb, kind="iou3d")
    rows = torch.arange(0, 60000, 997, device=ta.device)
    mat = U.boxes_iou3d_gpu(ta[rows], tb)
    w = mat.max(1)
    assert torch.equal(rmax[rows], w.values)
    assert torch.equal(rarg[rows], (mat == w.values.unsqueeze(1)).int().argmax(1))
    assert float((rmax > 0).float().mean()) > 0.4  # half of the second set are jittered copies of the first
