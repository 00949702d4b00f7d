"""Import the compiled UNMODIFIED reference extensions from oracle/_ref/ (built by oracle/build_ref.py).

Returns None when they are not there (a checkout that never ran build_ref.py): tests that need
them skip, everything else falls back to the restatement in lg_oracle.c.  TEST INFRASTRUCTURE ONLY.
"""
import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def _load(name):
    if name in _cache:
        return _cache[name]
    path = os.path.join(_HERE, "_ref", name, name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the extension links against libtorch)

        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    _cache[name] = mod
    return mod


def iou3d_nms_cuda():
    return _load("iou3d_nms_cuda")


def roiaware_pool3d_cuda():
    return _load("roiaware_pool3d_cuda")


def roipoint_pool3d_cuda():
    return _load("roipoint_pool3d_cuda")


def available():
    return iou3d_nms_cuda() is not None and roiaware_pool3d_cuda() is not None
