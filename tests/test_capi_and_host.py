"""CPU suite: the C-ABI library loads and exports every symbol include/tsmdet_b200.h declares (no
compute without a GPU), the host-side logic (BN folding, sharding maths, product isolation)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tsmdet_b200.h")
PKG = os.path.join(ROOT, "tsm-det-pointcloud-_b200")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsmdet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from tsmdet_b200 import _lib

    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the Python binding table covers the header
    assert set(names) <= set(_lib.EXPORTS), set(names) - set(_lib.EXPORTS)


def test_library_has_sm100a_code_and_no_torch_dependency():
    from tsmdet_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in ldd and "libc10" not in ldd


def test_version_and_error_strings():
    from tsmdet_b200 import _lib

    assert "sm_100a" in _lib.version()
    assert _lib.lib().tsmdet_error_string(0) == b"ok"
    assert b"invalid" in _lib.lib().tsmdet_error_string(1000001)
    with pytest.raises(_lib.TsmdetError):
        _lib.check(1000001, "unit-test")


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (a CPU fallback would void every parity claim)."""
    bad = []
    for dp, _, fs in os.walk(PKG):
        if os.path.basename(dp) == "build":
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\bimport\s+oracle\b|\bfrom\s+oracle\b|liboracle|oracle\.py", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_cuda_ops_refuse_cpu_tensors():
    from tsmdet_b200 import pointnet2_batch_cuda as ext

    with pytest.raises(ValueError):
        ext.ball_query_wrapper(1, 4, 1, 1.0, 2, torch.zeros(1, 1, 3), torch.zeros(1, 4, 3),
                               torch.zeros(1, 1, dtype=torch.int32), torch.zeros(1, 1, 2, dtype=torch.int32))


def test_fold_conv_bn_matches_eval_stack():
    from tsmdet_b200.pointnet2_modules import build_shared_mlp, fold_conv_bn

    torch.manual_seed(0)
    mlp = build_shared_mlp([7, 16, 12], bn=True).eval()
    for m in mlp.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    x = torch.randn(3, 7, 5, 4)
    with torch.no_grad():
        want = mlp(x)
        y = x
        for w, b in fold_conv_bn(mlp):
            y = torch.relu(torch.einsum("oc,bcpq->bopq", w, y) + b[None, :, None, None])
    assert torch.allclose(y, want, atol=1e-5)


def test_fps_plan_introspection():
    """Launch planning is host logic and runs without a GPU (SM count falls back to 148)."""
    from tsmdet_b200 import _lib

    c, t, p, s = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
    for b, n in [(16, 16384), (16, 4096), (16, 1024), (8, 65536), (1, 180000), (200, 16384), (1, 1), (3, 37)]:
        _lib.call("tsmdet_fps_plan", b, n, ctypes.byref(c), ctypes.byref(t), ctypes.byref(p), ctypes.byref(s))
        bs = 1 << int(np.log(float(n)) / np.log(2.0))
        bs = max(min(bs, 1024), 1)
        assert c.value in (1, 2, 4, 8, 16) and t.value in (128, 256, 512, 1024)
        assert c.value * t.value >= bs, "cluster threads must cover the reference's block size"
        assert c.value * t.value * p.value >= n, "every point needs a register slot"


def test_shard_bounds_cover_everything():
    from tsmdet_b200.sharding import shard_bounds

    for total in (0, 1, 7, 16, 128, 130):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_import_sets_hardware_queue_default_and_respects_an_existing_value():
    """`import tsmdet_b200` asks for 32 hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) before the process creates its
    CUDA context -- the in-kernel waits of the peer gather must not share a queue with other pipeline lanes --
    and leaves a value the user exported alone."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "import os, sys; sys.path.insert(0, %r); import tsmdet_b200; print(os.environ['CUDA_DEVICE_MAX_CONNECTIONS'])" % root
    env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip()
    assert out == "32"
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "4"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip()
    assert out == "4"
