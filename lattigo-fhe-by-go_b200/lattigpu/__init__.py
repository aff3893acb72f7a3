"""lattigpu -- B200-native RNS polynomial-ring engine behind Lattigo's ring.Context API.

Python mirror of the reference's Go interface, bound to the C ABI in
include/lattigpu.h.  GPU only: importing works anywhere, calling any op
without the built library or without a CUDA device raises.
"""
from . import bfv, bfv_scheme, ckks, ckks_scheme, dbfv, dckks, dist, marshaler, ring  # noqa: F401
from ._lib import LIB_PATH, LattigpuError, lib  # noqa: F401

__all__ = ["ring", "ckks", "ckks_scheme", "bfv", "bfv_scheme", "dckks", "dbfv", "dist", "marshaler", "lib", "LattigpuError", "LIB_PATH"]
