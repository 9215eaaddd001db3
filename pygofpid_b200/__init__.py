"""pygofpid_b200 -- B200-native (sm_100a) implementation of PyGOFPID's per-frame
foreground-extraction hot path behind the reference's `GOFPID` pipeline class.

The arithmetic lives in libgofpid_b200.so (hand-written CUDA, C ABI in include/gofpid_b200.h);
this package is the thin host side.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .hotpath import HotPath  # noqa: F401

__all__ = ["HotPath", "__version__"]
