"""ntjoin_b200 -- B200-native minimizer sketch-and-filter engine (drop-in for ntJoin steps 1-3).

Host side is Python (like the reference's bin/*.py); all arithmetic runs in hand-written sm_100a
CUDA kernels behind the C ABI of include/mxe.h (ntjoin_b200/libmxe.so).  There is no CPU fallback:
importing works anywhere, creating an Engine without a CUDA device or without the built library
raises.
"""
from ._lib import MxeError, load_library, library_path  # noqa: F401
from .engine import Engine, Sketch, FilterResult, HostSketch  # noqa: F401

__all__ = ["Engine", "Sketch", "FilterResult", "HostSketch", "MxeError", "load_library", "library_path"]
__version__ = "0.1.0"
