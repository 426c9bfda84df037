import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): builds liboracle.so on first use."""
    from oracle import oracle as _orc

    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def ref_pointnet2():
    """The UNMODIFIED reference CUDA extension (oracle/_ref), or None when it is not built."""
    from oracle import build_ref

    return build_ref.load_ref("pointnet2_batch_cuda")


@pytest.fixture(scope="session")
def ref_iou3d():
    from oracle import build_ref

    return build_ref.load_ref("iou3d_nms_cuda")


# ---- TSMDET_* tuning knobs: the library reads them from the environment ONCE (no getenv on launch paths), so a
# test that flips one has to ask for a re-read -- and the next test must not inherit it.
def knob_setenv(monkeypatch, name, value):
    from tsmdet_b200 import _lib

    monkeypatch.setenv(name, value)
    _lib.reload_options()


def knob_delenv(monkeypatch, name, raising=True):
    from tsmdet_b200 import _lib

    monkeypatch.delenv(name, raising=raising)
    _lib.reload_options()


@pytest.fixture(autouse=True)
def _fresh_knobs():
    """Autouse fixtures are set up before (and torn down after) ``monkeypatch``: the re-read in the teardown sees
    the restored environment."""
    yield
    try:
        from tsmdet_b200 import _lib
    except ImportError:
        return
    _lib.reload_options()
