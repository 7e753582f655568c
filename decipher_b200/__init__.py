"""decipher_b200 — B200-native Monte-Carlo rainfall-runoff path of DECIPHeR (hot path only).

The product is the CUDA shared library `csrc/build/libdecipher_b200.so` behind the C ABI of
`include/decipher_b200.h`; this package is the thin Python host around it (ctypes bindings, the
synthetic-catchment generator and the reference-format file readers/writers).
"""
from .capi import (Catchment, DeviceCatchment, DecipherError, load_library, measure_fp64_peak)  # noqa: F401

__all__ = ["Catchment", "DeviceCatchment", "DecipherError", "load_library", "measure_fp64_peak"]
