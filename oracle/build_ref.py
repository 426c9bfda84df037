"""Build recipe for ``oracle/_ref``: the UNMODIFIED reference CUDA extensions.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path.

The reference's hot-path native code compiles from its own few source files:

* ``/root/reference/pcdet/ops/pointnet2/pointnet2_batch/src/*.{cpp,cu}``
  -> ``oracle/_ref/pointnet2_batch_cuda.so``  (15 pybind functions,
  ``pointnet2_api.cpp:10-33``)
* ``/root/reference/pcdet/ops/iou3d_nms/src/*.{cpp,cu}``
  -> ``oracle/_ref/iou3d_nms_cuda.so``        (5 pybind functions,
  ``iou3d_nms_api.cpp:11-17``)
* ``/root/reference/pcdet/ops/pointnet2/pointnet2_stack/src/*.{cpp,cu}``
  -> ``oracle/_ref/pointnet2_stack_cuda.so``  (voxel query, stack grouping,
  stack FPS: SURVEY.md 8 f3)

The sources are compiled where they lie (read-only); only build products are
written, and only under ``oracle/_ref/`` (git-ignored, but NOT gpurun-ignored, so
the ``.so`` files travel to the GPU box).  Flags mirror the reference's
``setup.py:17-22`` (no extra compile args: nvcc default ``-O3``, ``-fmad=true``,
no fast-math) plus the one arch flag for B200.

The GPU kernels can only *run* on a CUDA device, so on the GPU box they are the
bit-exact oracle for indices / keep-lists and the "GPU bar to beat" in
``bench.py --impl reference``; ``boxes_iou_bev_cpu`` runs anywhere.
"""
from __future__ import annotations

import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TSMDET_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "pointnet2_batch_cuda": "pcdet/ops/pointnet2/pointnet2_batch/src",
    "iou3d_nms_cuda": "pcdet/ops/iou3d_nms/src",
    # voxel query (+dilated), stack grouping, stack FPS (15 pybind functions, src/pointnet2_api.cpp:12-32)
    "pointnet2_stack_cuda": "pcdet/ops/pointnet2/pointnet2_stack/src",
}


def built(name: str) -> str | None:
    p = os.path.join(OUT, name + ".so")
    return p if os.path.exists(p) else None


def build(verbose: bool = False) -> dict:
    """Compile both reference extensions; returns {name: path-or-None}."""
    res = {}
    if not os.path.isdir(REF_ROOT):
        return {k: built(k) for k in EXTS}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    for name, rel in EXTS.items():
        if built(name):
            res[name] = built(name)
            continue
        src_dir = os.path.join(REF_ROOT, rel)
        sources = sorted(glob.glob(os.path.join(src_dir, "*.cpp")) + glob.glob(os.path.join(src_dir, "*.cu")))
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(
            name=name,
            sources=sources,
            extra_include_paths=[src_dir],
            # host .cpp flags = what the reference's setup.py build gets from distutils
            # (sysconfig CFLAGS).  They matter: iou3d_cpu.cpp and iou3d_nms_kernel.cu
            # both define inline intersection()/box_overlap()/...; at -O0 the linker
            # resolves the CPU path to the .cu file's host stubs, which exit(1).
            extra_cflags=["-fno-strict-overflow", "-DNDEBUG", "-O2"],
            extra_cuda_cflags=["-gencode=arch=compute_100a,code=sm_100a"],
            build_directory=bdir,
            verbose=verbose,
            is_python_module=False,
        )
        so = os.path.join(bdir, name + ".so")
        dst = os.path.join(OUT, name + ".so")
        if os.path.exists(so):
            os.replace(so, dst)
        res[name] = built(name)
    return res


# The reference's Python op wrappers (the drop-in boundary's CALLERS), byte-compiled where they lie: the .pyc files
# are build products like the .so files above (no source is copied), they travel to the GPU box, and
# tests/test_dropin_reference_py_gpu.py runs them UNMODIFIED over this repo's pybind-name shims.
PY_FILES = {
    "pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils": "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py",
    "pcdet.ops.iou3d_nms.iou3d_nms_utils": "pcdet/ops/iou3d_nms/iou3d_nms_utils.py",
    "pcdet.models.model_utils.model_nms_utils": "pcdet/models/model_utils/model_nms_utils.py",
}


def pyc_path(modname: str) -> str:
    return os.path.join(OUT, "pyc", modname + ".pycbin")  # not *.pyc: snapshot tools tend to drop those


def build_py() -> dict:
    """Byte-compile the reference's three Python op files into oracle/_ref/pyc/ (sourceless, unmodified)."""
    import py_compile

    res = {}
    for mod, rel in PY_FILES.items():
        dst = pyc_path(mod)
        src = os.path.join(REF_ROOT, rel)
        if os.path.exists(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
        res[mod] = dst if os.path.exists(dst) else None
    return res


def load_ref(name: str):
    """Import a built reference extension as a Python module (or None)."""
    p = built(name)
    if p is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv)
    out.update(build_py())
    print(out)
    sys.exit(0 if all(out.values()) else 1)
