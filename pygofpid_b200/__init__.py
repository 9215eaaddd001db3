"""pygofpid_b200 -- B200-native (sm_100a) implementation of PyGOFPID's per-frame
foreground-extraction hot path behind the reference's `GOFPID` pipeline class.

The arithmetic lives in libgofpid_b200.so (hand-written CUDA, C ABI in include/gofpid_b200.h);
this package is the thin host side.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .hotpath import HotPath  # noqa: F401
from .pipeline import GOFPID  # noqa: F401
from .segmentation import ForegroundEstimator, FrameDifferencing, ViBe  # noqa: F401

__all__ = ["GOFPID", "HotPath", "ForegroundEstimator", "FrameDifferencing", "ViBe", "__version__"]
