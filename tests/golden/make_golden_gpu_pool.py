#!/usr/bin/env python
"""Generate golden_gpu_pool.npz from the UNMODIFIED reference CUDA kernels of RoI-aware pooling and RoI point pooling
(oracle/_ref, compiled for sm_100a by oracle/build_ref.py) running on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu_pool.py'      # writes gpurun_out/golden_gpu_pool.npz
    cp gpurun_out/golden_gpu_pool.npz tests/golden/

Inputs are regenerated from seeds by the tests (cases() below); only the reference's outputs are stored.  The reference
has no CPU implementation of these functions, so this file is what pins the oracle's restatement of them
(tests/test_oracle_pin.py) in the CPU-only suite.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidardetection_b200 import synth  # noqa: E402

# name: (n_points, n_rois, channels, out_size, max_pts_each_voxel, num_sampled_points, seed)
CASES = {
    "small": (2048, 8, 3, (4, 4, 4), 8, 64, 901),
    "ragged": (3001, 5, 7, (3, 5, 2), 4, 33, 902),
    "deep": (4096, 6, 16, (6, 6, 6), 128, 512, 903),
}


def cases():
    for name, (m, n, c, out, mp, s, seed) in CASES.items():
        pts, rois, feat = synth.pool_case(m, n, c, seed)
        rois = rois.copy()
        rois[0, 3:5] *= 4.0  # a big box: voxels overflow max_pts, more than S points inside
        # a box of zero length with points on its axis: x_res = 0, so (local_x + 0) / 0 is +-inf or nan and the
        # float -> int conversion saturates (the unsigned clamp then picks the last voxel or voxel 0)
        rois[-1, 3] = 0.0
        cx, cy, cz, _, dy, dz, rz = rois[-1]
        pts = pts.copy()
        for i, t in enumerate((-0.4, -0.1, 0.0, 0.2, 0.45)):
            pts[i] = (cx - np.sin(rz) * t * dy, cy + np.cos(rz) * t * dy, cz + 0.3 * t * dz)
        yield name, pts, rois, feat, out, mp, s


def main():
    import torch
    from oracle import ref_loader as R

    dev = torch.device("cuda:0")
    roi, rpp = R.roiaware_pool3d_cuda(), R.roipoint_pool3d_cuda()
    assert roi is not None and rpp is not None, "oracle/_ref missing (build it in the dev container first)"
    out = {"gpu_name": np.array(torch.cuda.get_device_name(0))}
    for name, pts, rois, feat, osz, mp, s in cases():
        tr, tp, tf = (torch.from_numpy(x).to(dev).contiguous() for x in (rois, pts, feat))
        n, c = rois.shape[0], feat.shape[1]
        for method, mname in ((0, "max"), (1, "avg")):
            # roiaware_pool3d_utils.py:84-90: three zero-filled outputs, then the extension call
            pooled = tf.new_zeros((n, *osz, c))
            argmax = tf.new_zeros((n, *osz, c), dtype=torch.int)
            pidx = tf.new_zeros((n, *osz, mp), dtype=torch.int)
            roi.forward(tr, tp, tf, argmax, pidx, pooled, method)
            g = torch.from_numpy(np.random.default_rng(7).standard_normal(tuple(pooled.shape)).astype(np.float32)).to(dev)
            grad_in = g.new_zeros((pts.shape[0], c))
            roi.backward(pidx, argmax, g.contiguous(), grad_in, method)
            torch.cuda.synchronize()
            out[f"{name}_{mname}_pooled"] = pooled.cpu().numpy()
            out[f"{name}_{mname}_grad_in"] = grad_in.cpu().numpy()
            if method == 0:
                out[f"{name}_argmax"] = argmax.cpu().numpy()
                out[f"{name}_pts_idx"] = pidx.cpu().numpy()
        # roipoint_pool3d_utils.py:53-61 with pool_extra_width (0.2, 0.2, 0.2) already applied
        big = rois.copy()
        big[:, 3:6] += np.float32(0.2)
        tb = torch.from_numpy(big[None]).to(dev).contiguous()
        pf = tf.new_zeros((1, n, s, 3 + c))
        flag = tf.new_zeros((1, n)).int()
        rpp.forward(tp[None].contiguous(), tb, tf[None].contiguous(), pf, flag)
        torch.cuda.synchronize()
        out[f"{name}_rp_pooled"] = pf.cpu().numpy()
        out[f"{name}_rp_flag"] = flag.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    p = os.path.join(ROOT, "gpurun_out", "golden_gpu_pool.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
