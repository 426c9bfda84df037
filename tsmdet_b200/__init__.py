"""Import alias: the product package lives in ``tsm-det-pointcloud-_b200/`` (a directory name
Python cannot import directly); ``import tsmdet_b200`` resolves its submodules from there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tsm-det-pointcloud-_b200")
__path__.insert(0, _PKG_DIR)

from ._version import __version__  # noqa: E402,F401
