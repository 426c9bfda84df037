"""Import alias: the product package lives in ``tsm-det-pointcloud-_b200/`` (a directory name
Python cannot import directly); ``import tsmdet_b200`` resolves its submodules from there."""
import os as _os

# Hardware work queues.  The pipelined runner keeps 8 steps in flight on separate streams, each a CUDA graph with three
# concurrent branches, and at N > 1 follows every replay with the peer-gather kernels, which WAIT inside the kernel (credits,
# arrival flags).  With the CUDA default of 8 connections those streams share hardware queues, and a waiting kernel at the
# head of a queue holds back unrelated lanes behind it: measured on 2 x B200, 63.3k -> 65.4k frames/s (= 2 x one GPU) with
# 32 connections.  Read by the driver when the CUDA context is created, so it must be set before the first CUDA call of
# the process -- import this package first, or export the variable yourself (an existing value is respected).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tsm-det-pointcloud-_b200")
__path__.insert(0, _PKG_DIR)

from ._version import __version__  # noqa: E402,F401
