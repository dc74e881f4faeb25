"""yume_b200 — B200-native (sm_100a) implementation of YUME's denoise hot path behind the reference's own seams.

Public surface:
  yume_b200.ops       tensor-level wrappers over the C ABI (include/yume_b200.h)
  yume_b200.build     in-tree nvcc build of csrc/libyume_b200.so
"""
from ._lib import YumeB200Error, lib_path, load  # noqa: F401

__version__ = "0.1.0"
