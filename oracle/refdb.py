"""refdb.py - TEST INFRASTRUCTURE: ctypes driver for oracle/_ref/libmidoridb_ref.so.

The shared object is the UNMODIFIED reference compiled in place by oracle/Makefile (target `ref`)
plus oracle/ref_shim.c.  It exposes the reference's public API (include/engine/query.h:42-69,
include/engine/database.h:26-32) and the refh_* helpers.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libmidoridb_ref.so")
# the same reference objects with executor_run_select_stmt redirected to this repo's CUDA backend (oracle/ref_gpu_bridge.c)
GPU_LIB_PATH = os.path.join(HERE, "_ref", "libmidoridb_ref_gpu.so")

# enum COLUMN_TYPE, include/primitive/column.h:17-25
CT_VARCHAR, CT_INTEGER, CT_TINYINT, CT_DOUBLE, CT_DATE, CT_DATETIME = range(6)
# enum query_output_status, include/engine/query.h:15-22
ST_OK_WITH_RESULTS, ST_OK_EXECUTED, ST_ERROR = range(3)
MIDORIDB_OK, MIDORIDB_ROW = 0, 4

_libs = {}


def available():
    return os.path.exists(LIB_PATH)


def gpu_bridge_available():
    return os.path.exists(GPU_LIB_PATH)


def lib(path=LIB_PATH):
    _lib = _libs.get(path)
    if _lib is None:
        L = C.CDLL(path)
        vp = C.c_void_p
        L.refh_database_new.restype = vp
        L.refh_database_free.argtypes = [vp]
        L.refh_table_create.restype = vp
        L.refh_table_create.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
        L.refh_table_get.restype = vp
        L.refh_table_get.argtypes = [vp, C.c_char_p]
        L.refh_table_append.restype = C.c_long
        L.refh_table_append.argtypes = [vp, C.c_size_t, vp, vp]
        L.refh_table_row_size.restype = C.c_size_t
        L.refh_table_row_size.argtypes = [vp]
        L.refh_table_ncols.restype = C.c_int
        L.refh_table_ncols.argtypes = [vp]
        L.refh_table_colname.restype = C.c_char_p
        L.refh_table_colname.argtypes = [vp, C.c_int]
        L.refh_table_coltype.restype = C.c_int
        L.refh_table_coltype.argtypes = [vp, C.c_int]
        L.refh_table_npages.restype = C.c_size_t
        L.refh_table_npages.argtypes = [vp]
        L.refh_table_pages.restype = C.c_size_t
        L.refh_table_pages.argtypes = [vp, C.POINTER(vp), C.c_size_t]
        L.refh_table_delete_slot.restype = C.c_int
        L.refh_table_delete_slot.argtypes = [vp, C.c_size_t, C.c_size_t]
        L.refh_table_dump.restype = C.c_size_t
        L.refh_table_dump.argtypes = [vp, vp, vp, C.c_size_t]
        L.query_execute.restype = vp
        L.query_execute.argtypes = [vp, C.c_char_p]
        L.query_free.argtypes = [vp]
        L.query_cur_step.restype = C.c_int
        L.query_cur_step.argtypes = [vp]
        L.query_column_int64.restype = C.c_int64
        L.query_column_int64.argtypes = [vp, C.c_int]
        L.refh_output_status.restype = C.c_int
        L.refh_output_status.argtypes = [vp]
        L.refh_output_error.restype = C.c_char_p
        L.refh_output_error.argtypes = [vp]
        L.refh_output_rows_affected.restype = C.c_size_t
        L.refh_output_rows_affected.argtypes = [vp]
        L.refh_output_table.restype = vp
        L.refh_output_table.argtypes = [vp]
        L.refh_output_results.restype = vp
        L.refh_output_results.argtypes = [vp]
        if path == GPU_LIB_PATH:
            L.refh_select_backend.restype = C.c_char_p
        _libs[path] = _lib = L
    return _lib


class RefResult:
    """Materialised result of a reference SELECT: column names, raw int64 cells, null flags."""

    def __init__(self, names, types, cells, nulls, cursor_rows):
        self.names, self.types, self.cells, self.nulls, self.cursor_rows = names, types, cells, nulls, cursor_rows

    def rows(self):
        """rows as tuples with None for NULL (raw 8-byte cells as int64)"""
        out = []
        for r in range(self.cells.shape[0]):
            out.append(tuple(None if self.nulls[r, c] else int(self.cells[r, c]) for c in range(self.cells.shape[1])))
        return out


class RefDatabase:
    """The reference engine behind its own public C API."""

    def __init__(self, gpu_bridge=False):
        """gpu_bridge: the reference with its SELECT executor replaced by libmidoridb_cuda.so (oracle/ref_gpu_bridge.c)"""
        self.L = lib(GPU_LIB_PATH if gpu_bridge else LIB_PATH)
        self.db = self.L.refh_database_new()
        if not self.db:
            raise RuntimeError("database_open failed")

    def close(self):
        if self.db:
            self.L.refh_database_free(self.db)
            self.db = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def create_table(self, name, col_names, col_types):
        n = len(col_names)
        names = (C.c_char_p * n)(*[c.encode() for c in col_names])
        types = (C.c_int * n)(*col_types)
        t = self.L.refh_table_create(self.db, name.encode(), n, names, types)
        if not t:
            raise RuntimeError("refh_table_create failed for %s" % name)
        return t

    def table(self, name):
        return self.L.refh_table_get(self.db, name.encode())

    def append(self, table, cells, nulls=None):
        cells = np.ascontiguousarray(cells)
        if cells.dtype != np.int64:
            cells = cells.view(np.int64) if cells.dtype == np.float64 else cells.astype(np.int64)
        if cells.ndim == 1:
            cells = cells.reshape(-1, 1)
        nptr = None
        if nulls is not None:
            nulls = np.ascontiguousarray(nulls, dtype=np.uint8).reshape(cells.shape)
            nptr = nulls.ctypes.data
        rc = self.L.refh_table_append(table, cells.shape[0], cells.ctypes.data, nptr)
        if rc != cells.shape[0]:
            raise RuntimeError("refh_table_append failed")

    def page_ptrs(self, table):
        n = self.L.refh_table_npages(table)
        arr = (C.c_void_p * max(n, 1))()
        self.L.refh_table_pages(table, arr, n)
        return [arr[i] for i in range(n)], self.L.refh_table_row_size(table)

    def dump_table(self, table):
        ncols = self.L.refh_table_ncols(table)
        n = self.L.refh_table_dump(table, None, None, 0)
        cells = np.zeros((n, ncols), dtype=np.int64)
        nulls = np.zeros((n, ncols), dtype=np.uint8)
        self.L.refh_table_dump(table, cells.ctypes.data, nulls.ctypes.data, n)
        names = [self.L.refh_table_colname(table, i).decode() for i in range(ncols)]
        types = [self.L.refh_table_coltype(table, i) for i in range(ncols)]
        return names, types, cells, nulls

    def execute(self, sql):
        """run a non-SELECT statement; returns rows affected"""
        out = self.L.query_execute(self.db, sql.encode())
        st = self.L.refh_output_status(out)
        err = self.L.refh_output_error(out).decode(errors="replace")
        aff = self.L.refh_output_rows_affected(out)
        self.L.query_free(out)
        if st != ST_OK_EXECUTED:
            raise RuntimeError("reference rejected %r: %s" % (sql, err))
        return aff

    def query(self, sql, use_cursor=False):
        """run a SELECT; result read by page walk (and optionally also through the cursor API)"""
        out = self.L.query_execute(self.db, sql.encode())
        st = self.L.refh_output_status(out)
        if st != ST_OK_WITH_RESULTS:
            err = self.L.refh_output_error(out).decode(errors="replace")
            self.L.query_free(out)
            raise RuntimeError("reference rejected %r: %s" % (sql, err))
        table = self.L.refh_output_table(out)
        names, types, cells, nulls = self.dump_table(table)
        cursor_rows = None
        if use_cursor:
            rs = self.L.refh_output_results(out)
            cursor_rows = []
            while self.L.query_cur_step(rs) == MIDORIDB_ROW:
                cursor_rows.append(tuple(self.L.query_column_int64(rs, c) for c in range(len(names))))
        self.L.query_free(out)
        return RefResult(names, types, cells, nulls, cursor_rows)
