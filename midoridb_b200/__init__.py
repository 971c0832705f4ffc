"""midoridb_b200 - B200-native execution backend for MidoriDB's scan -> WHERE -> INNER JOIN -> GROUP BY path.

Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/midoridb_cuda.h) -> libmidoridb_cuda.so
  host/   host side mirroring the reference's public C API (query_execute ...) on top of the C ABI
  capi.py ctypes binding of the C ABI (tests, bench)
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
