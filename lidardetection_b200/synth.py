"""Seeded synthetic KITTI / Waymo / NuScenes-shaped inputs for the tests and bench.py.

All outputs are float32 numpy arrays; boxes are (N, 7) [x, y, z, dx, dy, dz, heading], points (M, 3).
There is no network and no dataset in this environment: shapes and statistics follow the reference's
configs (tools/cfgs/kitti_models/pointpillar.yaml:81-112, second.yaml:94-99, pv_rcnn.yaml:168-227,
nuscenes_models/cbgs_second_multihead.yaml:196-206, anchor_generator.py:67-74).
"""
import numpy as np

SEEDS = {"cfg1": 101, "cfg2": 202, "cfg3": 303, "cfg4": 404, "cfg5": 505, "dense": 606}

KITTI_PRIORS = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], dtype=np.float32)
KITTI_PRIOR_P = np.array([0.60, 0.25, 0.15])
WAYMO_PRIORS = np.array([[4.7, 2.1, 1.7], [0.91, 0.86, 1.73], [1.78, 0.84, 1.78]], dtype=np.float32)


def _rng(seed):
    return np.random.default_rng(seed)


def kitti_anchors():
    """321,408 anchors exactly as AnchorGenerator lays them out for pointpillar.yaml:81-112:
    z-major (1) x y(248) x x(216) x class(3)... flattened to (216*248*3*2, 7)."""
    xs = np.linspace(0.0, 69.12, 216, dtype=np.float32)
    ys = np.linspace(-39.68, 39.68, 248, dtype=np.float32)
    zc = np.array([-1.0, 0.265, 0.265], dtype=np.float32)  # anchor_bottom_heights + dz/2
    rots = np.array([0.0, 1.57], dtype=np.float32)
    out = np.empty((248, 216, 3, 2, 7), dtype=np.float32)
    out[..., 0] = xs[None, :, None, None]
    out[..., 1] = ys[:, None, None, None]
    out[..., 2] = zc[None, None, :, None]
    out[..., 3:6] = KITTI_PRIORS[None, None, :, None, :]
    out[..., 6] = rots[None, None, None, :]
    return out.reshape(-1, 7)


def gt_boxes(n, seed, x_range=(0.0, 69.0), y_range=(-39.0, 39.0), priors=KITTI_PRIORS, prior_p=KITTI_PRIOR_P, z=-1.0):
    r = _rng(seed)
    cls = r.choice(len(priors), size=n, p=prior_p if prior_p is not None else None)
    b = np.empty((n, 7), dtype=np.float32)
    b[:, 0] = r.uniform(*x_range, n)
    b[:, 1] = r.uniform(*y_range, n)
    b[:, 2] = z
    b[:, 3:6] = priors[cls] * (1.0 + 0.1 * r.standard_normal((n, 3))).astype(np.float32)
    b[:, 3:6] = np.maximum(b[:, 3:6], 0.05)
    b[:, 6] = r.uniform(-np.pi, np.pi, n)
    return b


def jitter_boxes(objs, n, r, pos_sigma=0.25, dim_sigma=0.05, rot_sigma=0.1, flip_p=0.05):
    """n detections scattered around `objs` (cfg2 recipe): centre noise proportional to size,
    dims * (1 + 5%), heading + N(0, 0.1), +pi/2 with probability 5%."""
    k = r.integers(0, len(objs), n)
    o = objs[k]
    d = o.copy()
    d[:, 0] += (r.standard_normal(n) * pos_sigma * o[:, 3]).astype(np.float32)
    d[:, 1] += (r.standard_normal(n) * pos_sigma * o[:, 4]).astype(np.float32)
    d[:, 2] += (r.standard_normal(n) * 0.1).astype(np.float32)
    d[:, 3:6] *= (1.0 + dim_sigma * r.standard_normal((n, 3))).astype(np.float32)
    d[:, 3:6] = np.maximum(d[:, 3:6], 0.05)
    d[:, 6] += (r.standard_normal(n) * rot_sigma).astype(np.float32)
    d[:, 6] += np.where(r.random(n) < flip_p, np.float32(np.pi / 2), np.float32(0)).astype(np.float32)
    return d.astype(np.float32)


def nms_frames(n_frames, n_boxes, seed, k_range=(10, 60), xy=((0.0, 69.0), (-39.0, 39.0))):
    """cfg2 / cfg5: per frame K objects, n_boxes clustered detections, distinct scores in (0.1, 1)."""
    r = _rng(seed)
    boxes = np.empty((n_frames, n_boxes, 7), dtype=np.float32)
    scores = np.empty((n_frames, n_boxes), dtype=np.float32)
    for f in range(n_frames):
        k = int(r.integers(k_range[0], k_range[1] + 1))
        objs = gt_boxes(k, int(r.integers(1 << 30)), x_range=xy[0], y_range=xy[1])
        boxes[f] = jitter_boxes(objs, n_boxes, r)
        # distinct scores: a random permutation of an evenly spaced ladder (spacing >> fp32 ulp)
        ladder = np.linspace(0.1, 1.0, n_boxes, endpoint=False, dtype=np.float64)
        scores[f] = r.permutation(ladder).astype(np.float32)
    return boxes, scores


def cfg1(seed=SEEDS["cfg1"]):
    """PointPillars KITTI anchor-target IoU: 321,408 anchors x 20 GT."""
    return kitti_anchors(), gt_boxes(20, seed)


def cfg2(n_frames=64, n_boxes=4096, seed=SEEDS["cfg2"]):
    """SECOND KITTI post-processing NMS: 4096 boxes / frame, thresh 0.01, 64 frames."""
    return nms_frames(n_frames, n_boxes, seed)


def cfg3(n_frames=1, n_points=16384, n_rois=100, seed=SEEDS["cfg3"]):
    """PV-RCNN: points (B, 16384, 3) and ROIs (B, 100, 7); 30% of the points fall inside a random ROI."""
    r = _rng(seed)
    pts = np.empty((n_frames, n_points, 3), dtype=np.float32)
    rois = np.empty((n_frames, n_rois, 7), dtype=np.float32)
    for f in range(n_frames):
        b = gt_boxes(n_rois, int(r.integers(1 << 30)))
        rois[f] = b
        p = np.stack([r.uniform(0, 70.4, n_points), r.uniform(-40, 40, n_points), r.uniform(-3, 1, n_points)], 1)
        n_in = int(0.3 * n_points)
        k = r.integers(0, n_rois, n_in)
        loc = r.uniform(-0.5, 0.5, (n_in, 3)) * b[k, 3:6]
        c, s = np.cos(b[k, 6]), np.sin(b[k, 6])
        p[:n_in, 0] = b[k, 0] + loc[:, 0] * c - loc[:, 1] * s
        p[:n_in, 1] = b[k, 1] + loc[:, 0] * s + loc[:, 1] * c
        p[:n_in, 2] = b[k, 2] + loc[:, 2]
        pts[f] = p[r.permutation(n_points)].astype(np.float32)
    return pts, rois


def cfg3_iou(seed=SEEDS["cfg3"] + 1):
    """PV-RCNN ROI target assignment: 512 ROIs (jittered GT) x 20 GT."""
    r = _rng(seed)
    gt = gt_boxes(20, int(r.integers(1 << 30)))
    return jitter_boxes(gt, 512, r), gt


def cfg4(n=200_000, seed=SEEDS["cfg4"]):
    """Waymo-scale evaluation IoU: n + n boxes over +-75.2 m; second set = 50% jittered first set + 50% fresh."""
    r = _rng(seed)

    def fresh(m):
        b = gt_boxes(m, int(r.integers(1 << 30)), x_range=(-75.2, 75.2), y_range=(-75.2, 75.2), priors=WAYMO_PRIORS,
                     prior_p=np.array([0.6, 0.25, 0.15]))
        b[:, 2] = r.uniform(-2, 4, m)
        return b

    a = fresh(n)
    half = n // 2
    pick = r.permutation(n)[:half]
    b1 = jitter_boxes(a[pick], half, r)
    b = np.concatenate([b1, fresh(n - half)], 0)
    return a, b[r.permutation(n)].astype(np.float32)


def cfg5(n_frames=256, n_classes=10, n_boxes=1000, seed=SEEDS["cfg5"]):
    """NuScenes CBGS multi-head NMS: frames x classes problems of 1000 boxes, thresh 0.2."""
    b, s = nms_frames(n_frames * n_classes, n_boxes, seed, k_range=(5, 40), xy=((-51.2, 51.2), (-51.2, 51.2)))
    return b.reshape(n_frames, n_classes, n_boxes, 7), s.reshape(n_frames, n_classes, n_boxes)


def dense_overlap(n=16384, m=16384, seed=SEEDS["dense"], centre=(0.0, 0.0)):
    """FP32-roofline microbench: every pair overlaps (centres ~N(0, 0.3 m), car-sized, random heading)."""
    r = _rng(seed)

    def mk(k):
        b = np.empty((k, 7), dtype=np.float32)
        b[:, 0] = centre[0] + 0.3 * r.standard_normal(k)
        b[:, 1] = centre[1] + 0.3 * r.standard_normal(k)
        b[:, 2] = 0.0
        b[:, 3:6] = np.array([3.9, 1.6, 1.5], dtype=np.float32)
        b[:, 6] = r.uniform(-np.pi, np.pi, k)
        return b

    return mk(n), mk(m)


def clustered_pairs(n, m, seed, centre=(35.0, 17.5), priors=KITTI_PRIORS, sigma=0.3):
    """Differential-test workload (SURVEY App. B probe): n x m boxes clustered around `centre` with
    sigma = 0.3 * size, dims +-5%, heading U(-pi, pi) -- a large fraction of pairs overlap."""
    r = _rng(seed)

    def mk(k):
        cls = r.integers(0, len(priors), k)
        b = np.empty((k, 7), dtype=np.float32)
        b[:, 3:6] = priors[cls] * (1.0 + 0.05 * r.standard_normal((k, 3)))
        b[:, 0] = centre[0] + sigma * b[:, 3] * r.standard_normal(k)
        b[:, 1] = centre[1] + sigma * b[:, 4] * r.standard_normal(k)
        b[:, 2] = -1.0 + 0.2 * r.standard_normal(k)
        b[:, 6] = r.uniform(-np.pi, np.pi, k)
        return b.astype(np.float32)

    return mk(n), mk(m)


def pool_case(n_points=16384, n_rois=128, channels=128, seed=SEEDS["cfg3"] + 7):
    """Part-A2 / PointRCNN RoI pooling: one frame of cfg3-shaped points and ROIs plus (n_points, C) features."""
    pts, rois = cfg3(1, n_points, n_rois, seed)
    feat = _rng(seed + 1).standard_normal((n_points, channels)).astype(np.float32)
    return pts[0], rois[0], feat


def kitti_eval_frames(n_frames, seed, gt_range=(2, 12), fp_range=(2, 14)):
    """KITTI-evaluation shaped annotations in CAMERA coordinates (kitti_object_eval_python/eval.py:340-414 consumes them):
    per frame float64 boxes (x, y, z, l, h, w, rotation_y) with x right, y down (bottom face), z forward.
    Ground truth is rounded to 2 decimals (label_2 text files), detections to 4 (the reference's result writer);
    detections = jittered ground truth (90 % recall) + false positives.  Returns (list of gt (G_f,7), list of dt (D_f,7))."""
    r = _rng(seed)
    pri = KITTI_PRIORS.astype(np.float64)  # (l, w, h) lidar order
    gts, dts = [], []
    for _ in range(n_frames):
        g = int(r.integers(gt_range[0], gt_range[1] + 1))
        cls = r.choice(3, size=g, p=KITTI_PRIOR_P)
        dims = pri[cls] * (1.0 + 0.1 * r.standard_normal((g, 3)))
        dims = np.maximum(dims, 0.3)
        gt = np.empty((g, 7))
        gt[:, 0] = r.uniform(-35.0, 35.0, g)
        gt[:, 1] = r.normal(1.65, 0.25, g)
        gt[:, 2] = r.uniform(3.0, 68.0, g)
        gt[:, 3] = dims[:, 0]
        gt[:, 4] = dims[:, 2]
        gt[:, 5] = dims[:, 1]
        gt[:, 6] = r.uniform(-np.pi, np.pi, g)
        gt = np.round(gt, 2)
        hit = gt[r.random(g) < 0.9]
        det = hit.copy()
        n = len(det)
        det[:, 0] += r.normal(0, 0.15, n)
        det[:, 1] += r.normal(0, 0.05, n)
        det[:, 2] += r.normal(0, 0.15, n)
        det[:, 3:6] *= 1.0 + 0.05 * r.standard_normal((n, 3))
        det[:, 6] += r.normal(0, 0.08, n)
        f = int(r.integers(fp_range[0], fp_range[1] + 1))
        cls = r.choice(3, size=f, p=KITTI_PRIOR_P)
        fdim = pri[cls] * (1.0 + 0.1 * r.standard_normal((f, 3)))
        fp = np.empty((f, 7))
        fp[:, 0] = r.uniform(-35.0, 35.0, f)
        fp[:, 1] = r.normal(1.65, 0.25, f)
        fp[:, 2] = r.uniform(3.0, 68.0, f)
        fp[:, 3] = fdim[:, 0]
        fp[:, 4] = fdim[:, 2]
        fp[:, 5] = fdim[:, 1]
        fp[:, 6] = r.uniform(-np.pi, np.pi, f)
        dt = np.round(np.concatenate([det, fp], 0), 4)
        gts.append(gt)
        dts.append(dt[r.permutation(len(dt))])
    return gts, dts
