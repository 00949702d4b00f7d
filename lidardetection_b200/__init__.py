"""lidardetection_b200 -- B200-native rotated 3D-box geometry ops behind the reference's operator API.

Drop-in modules (same names / signatures / return types as zhengjingsen/LidarDetection's pcdet.ops):
    lidardetection_b200.ops.iou3d_nms.iou3d_nms_utils          <- pcdet/ops/iou3d_nms/iou3d_nms_utils.py
    lidardetection_b200.ops.roiaware_pool3d.roiaware_pool3d_utils  <- pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py
Everything is computed by liblidargeom.so (hand-written sm_100a CUDA behind a C ABI, include/lidargeom.h);
there is no CPU fallback: importing the ops without the built library raises.
"""
__version__ = "0.1.0"
