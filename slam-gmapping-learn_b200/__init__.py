"""B200-native GMapping per-scan particle update.

The product is the C-ABI CUDA library (include/gm_b200.h, csrc/) and the C++ host layer that keeps the
openslam_gmapping API (host/).  This Python package only binds the C ABI with ctypes for the tests and
the benchmark; it holds no compute path of its own and no CPU fallback.
"""
from .capi import GmContext, GmError, MatchParams, load_library  # noqa: F401
