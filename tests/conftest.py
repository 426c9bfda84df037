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
