"""Loads the reference's UNMODIFIED Python op files (byte-compiled by oracle/build_ref.py into oracle/_ref/pyc/,
test infrastructure) over THIS repo's pybind-name shims, exactly as INTEGRATION.md section 1 tells a maintainer to
wire them: ``pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda`` and ``pcdet.ops.iou3d_nms.iou3d_nms_cuda``
resolve to ``tsmdet_b200.pointnet2_batch_cuda`` / ``tsmdet_b200.iou3d_nms_cuda``; the ``pcdet`` parents are empty
stubs (the three files need nothing else but ``pcdet.utils.common_utils.check_numpy_to_torch``)."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub(name: str):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # a package
        sys.modules[name] = mod
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_stub(parent), leaf, mod)
    return mod


def _alias(name: str, mod):
    parent, _, leaf = name.rpartition(".")
    sys.modules[name] = mod
    setattr(_stub(parent), leaf, mod)


def _load_pyc(name: str):
    from oracle import build_ref

    path = build_ref.pyc_path(name)
    if not os.path.exists(path):
        return None
    if name in sys.modules:
        return sys.modules[name]
    parent, _, leaf = name.rpartition(".")
    _stub(parent)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = parent
    sys.modules[name] = mod
    loader.exec_module(mod)
    setattr(sys.modules[parent], leaf, mod)
    return mod


def load_reference_python():
    """Returns (pointnet2_utils, iou3d_nms_utils, model_nms_utils) of the reference, or None when oracle/_ref/pyc is
    not built."""
    from tsmdet_b200 import common_utils, iou3d_nms_cuda, pointnet2_batch_cuda

    _alias("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda", pointnet2_batch_cuda)
    _alias("pcdet.ops.iou3d_nms.iou3d_nms_cuda", iou3d_nms_cuda)
    _alias("pcdet.utils.common_utils", common_utils)
    mods = [_load_pyc(n) for n in ("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils",
                                   "pcdet.ops.iou3d_nms.iou3d_nms_utils",
                                   "pcdet.models.model_utils.model_nms_utils")]
    return None if any(m is None for m in mods) else tuple(mods)
