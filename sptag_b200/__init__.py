"""sptag_b200 -- B200-native drop-in for SPTAG's batched in-memory search path.

The product is the C-ABI shared library ``sptag_b200/lib/libsptag_b200.so`` (declared in
``include/sptag_b200.h``, built from ``sptag_b200/csrc``).  This package only holds the ctypes
binding used by the tests and ``bench.py``.
"""
from .capi import B200Index, SptagB200Error, launch_count, lib, merge_topk  # noqa: F401
