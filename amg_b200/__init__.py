"""amg_b200: B200-native (sm_100a) solve phase of the txthpc/amg classical AMG solver.

The product is amg_b200/libamgb200.so (C ABI in include/amg_b200.h); this package is the thin
ctypes host mirror used by tests/ and bench.py.  There is no CPU fallback.
"""
from . import capi  # noqa: F401
from .host import (CsrMatrix, DeviceHierarchy, HostHierarchy, generate, read_mtx,  # noqa: F401
                   read_mtx_fast, solve_dropin)
