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


def reference_python(module, ext, ext_name, package):
    """The reference's OWN Python wrapper module (`iou3d_nms_utils` / `roiaware_pool3d_utils`, byte-compiled unmodified by
    oracle/build_ref.py into oracle/_ref/py/) imported inside a synthetic package `package`, with `ext` standing where the
    module does `from . import <ext_name>` and a 4-line stub for `...utils.common_utils` (the real one drags in the whole
    detector framework).  Returns None when the bytecode is not there."""
    import sys
    import types

    path = os.path.join(_HERE, "_ref", "py", module + ".pyc.bin")
    if not os.path.exists(path):
        return None
    sub = {"iou3d_nms_utils": "iou3d_nms", "roiaware_pool3d_utils": "roiaware_pool3d"}[module]
    for name in (package, package + ".ops", package + ".utils", f"{package}.ops.{sub}"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    cu = types.ModuleType(package + ".utils.common_utils")

    def check_numpy_to_torch(x):  # pcdet/utils/common_utils.py:46-49
        import numpy as np
        import torch

        if isinstance(x, np.ndarray):
            return torch.from_numpy(x).float(), True
        return x, False

    cu.check_numpy_to_torch = check_numpy_to_torch
    sys.modules[package + ".utils.common_utils"] = cu
    sys.modules[package + ".utils"].common_utils = cu
    sys.modules[f"{package}.ops.{sub}.{ext_name}"] = ext
    setattr(sys.modules[f"{package}.ops.{sub}"], ext_name, ext)
    full = f"{package}.ops.{sub}.{module}"
    loader = importlib.machinery.SourcelessFileLoader(full, path)
    spec = importlib.util.spec_from_loader(full, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    loader.exec_module(mod)
    return mod
