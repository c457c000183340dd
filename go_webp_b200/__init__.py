"""go_webp_b200 -- B200 (sm_100a) implementation of go-webp's pixel hot path.

Layout:
  csrc/       hand-written CUDA kernels + the C ABI (include/b2webp.h)
  lib/        libb2webp.so (built in-tree by __graft_entry__.build())
  _abi.py     ctypes declarations of the C ABI
  device.py   contexts, device / pinned buffers, streams
  picture.py  mirror of pkg/picture + pkg/libwebp/enc entry points (parts 1, 2)
  dsp.py      mirror of pkg/libwebp/dsp function tables (parts 2, 3)
  partition.py  batch -> GPU sharding (no collective on the data path)

The CUDA library is the only implementation: importing a submodule that needs
it raises ImportError when libb2webp.so has not been built.
"""
from . import _abi  # noqa: F401

__all__ = ["_abi", "device", "picture", "dsp", "partition"]
