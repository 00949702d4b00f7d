"""Route B of INTEGRATION.md: keep the reference's OWN Python (pcdet/ops/iou3d_nms/iou3d_nms_utils.py,
pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py) and put ctypes shims with the extension modules' names and function
signatures where its compiled pybind extensions would be imported from:

    iou3d_nms_cuda        <- pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17
    roiaware_pool3d_cuda  <- pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177

`install("pcdet")` registers them in sys.modules as `pcdet.ops.iou3d_nms.iou3d_nms_cuda` and
`pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda`, so `from . import iou3d_nms_cuda` in the unmodified reference module
resolves to the shim (tests/test_route_b.py loads the reference's own files that way).
"""
import sys


def install(package="pcdet"):
    """Make `<package>.ops.iou3d_nms.iou3d_nms_cuda` and `<package>.ops.roiaware_pool3d.roiaware_pool3d_cuda` resolve to
    the shims.  Call before importing `<package>.ops.iou3d_nms.iou3d_nms_utils`.  Returns the two modules."""
    from . import iou3d_nms_cuda, roiaware_pool3d_cuda

    sys.modules[f"{package}.ops.iou3d_nms.iou3d_nms_cuda"] = iou3d_nms_cuda
    sys.modules[f"{package}.ops.roiaware_pool3d.roiaware_pool3d_cuda"] = roiaware_pool3d_cuda
    return iou3d_nms_cuda, roiaware_pool3d_cuda
