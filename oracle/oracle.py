"""oracle.py - TEST INFRASTRUCTURE: ctypes driver for oracle/libmdb_oracle.so (oracle/mdb_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.  It takes the same
`struct mdbcu_plan` as the CUDA library (built with midoridb_b200.capi.make_plan) with oracle table
handles in plan.tables[].
"""
import ctypes as C
import os
import subprocess

import numpy as np

from midoridb_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmdb_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.orc_last_error.restype = C.c_char_p
        L.orc_table_create.restype = vp
        L.orc_table_create.argtypes = [C.c_int, C.POINTER(C.c_int32)]
        L.orc_table_free.argtypes = [vp]
        L.orc_table_free.restype = None
        L.orc_table_rows.argtypes = [vp]
        L.orc_table_rows.restype = sz
        L.orc_table_append_pages.argtypes = [vp, vp, sz, sz]
        L.orc_table_append_columns.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(vp)]
        L.orc_select.argtypes = [C.POINTER(capi.Plan), C.POINTER(vp)]
        L.orc_result_rows.argtypes = [vp]
        L.orc_result_rows.restype = sz
        L.orc_result_cols.argtypes = [vp]
        L.orc_result_col_type.argtypes = [vp, C.c_int]
        L.orc_result_fetch_columns.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.orc_result_free.argtypes = [vp]
        L.orc_result_free.restype = None
        L.orc_join_count_groups.restype = C.c_int64
        L.orc_join_count_groups.argtypes = [vp, sz, vp, sz, vp, vp, sz]
        _lib = L
    return _lib


class OracleTable:
    def __init__(self, col_types):
        self.L = lib()
        self.col_types = list(col_types)
        types = (C.c_int32 * len(col_types))(*col_types)
        self.handle = self.L.orc_table_create(len(col_types), types)
        if not self.handle:
            raise RuntimeError("orc_table_create failed")

    def append_pages(self, pages, stride=capi.PAGE_SIZE):
        pages = np.ascontiguousarray(pages, dtype=np.uint8)
        rc = self.L.orc_table_append_pages(self.handle, pages.ctypes.data, pages.size // stride, stride)
        assert rc == 0

    def append_columns(self, columns, nulls=None):
        cols = [np.ascontiguousarray(c) for c in columns]
        n = cols[0].size
        data = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        nl, keep = None, []
        if nulls is not None:
            keep = [None if x is None else np.ascontiguousarray(x, dtype=np.uint8) for x in nulls]
            nl = (C.c_void_p * len(cols))(*[None if x is None else x.ctypes.data for x in keep])
        rc = self.L.orc_table_append_columns(self.handle, n, data, nl)
        assert rc == 0

    @property
    def rows(self):
        return self.L.orc_table_rows(self.handle)

    def free(self):
        if self.handle:
            self.L.orc_table_free(self.handle)
            self.handle = None


def select(plan):
    """run the oracle; returns (types, cells[list of arrays], nulls[list of arrays]) in the reference's row order"""
    L = lib()
    h = C.c_void_p()
    rc = L.orc_select(C.byref(plan), C.byref(h))
    if rc != 0:
        raise RuntimeError("oracle failed (%d): %s" % (rc, L.orc_last_error().decode()))
    n, ncols = L.orc_result_rows(h), L.orc_result_cols(h)
    types = [L.orc_result_col_type(h, c) for c in range(ncols)]
    cells = [np.zeros(n, dtype=np.float64 if t == capi.CT_DOUBLE else np.int64) for t in types]
    nulls = [np.zeros(n, dtype=np.uint8) for _ in types]
    cp = (C.c_void_p * max(ncols, 1))(*[c.ctypes.data for c in cells])
    npn = (C.c_void_p * max(ncols, 1))(*[x.ctypes.data for x in nulls])
    L.orc_result_fetch_columns(h, cp, npn)
    L.orc_result_free(h)
    return types, cells, nulls


def rows_of(cells, nulls):
    n = cells[0].size if cells else 0
    return [tuple(None if nulls[c][r] else cells[c][r].item() for c in range(len(cells))) for r in range(n)]


def join_count_groups(a, b, want_output=False):
    """single-thread hash join + count over raw int64 key arrays (cpu_baseline "port" for large sizes)"""
    L = lib()
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    if want_output:
        cap = min(a.size, b.size)
        keys = np.zeros(cap, dtype=np.int64)
        cnts = np.zeros(cap, dtype=np.int64)
        g = L.orc_join_count_groups(a.ctypes.data, a.size, b.ctypes.data, b.size, keys.ctypes.data, cnts.ctypes.data, cap)
        return keys[:g], cnts[:g]
    return L.orc_join_count_groups(a.ctypes.data, a.size, b.ctypes.data, b.size, None, None, 0)
