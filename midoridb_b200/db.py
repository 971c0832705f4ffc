"""ctypes binding of libmidoridb_b200.so (include/midoridb.h): MidoriDB's own public C API -
database_open / query_execute / query_cur_step / query_column_int64 / query_free / database_close
(reference: include/engine/query.h:42-69, include/engine/database.h:26-32) - served by the B200 backend.

The structures below mirror the C layouts so Python reads `struct query_output` the way a C caller would.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmidoridb_b200.so")

MIDORIDB_OK, MIDORIDB_ROW = 0, 4
ST_OK_WITH_RESULTS, ST_OK_EXECUTED, ST_ERROR = 0, 1, 2
CT_VARCHAR, CT_INTEGER, CT_TINYINT, CT_DOUBLE, CT_DATE, CT_DATETIME = range(6)


class ListHead(C.Structure):
    pass


ListHead._fields_ = [("next", C.POINTER(ListHead)), ("prev", C.POINTER(ListHead))]


class Column(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("type", C.c_int), ("precision", C.c_int), ("indexed", C.c_bool),
                ("nullable", C.c_bool), ("unique", C.c_bool), ("auto_inc", C.c_bool), ("primary_key", C.c_bool),
                ("is_count", C.c_bool)]


class Table(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("columns", Column * 128), ("column_count", C.c_int),
                ("datablock_head", C.POINTER(ListHead)), ("free_dtbkl_offset", C.c_size_t), ("mutex", C.c_ubyte * 40)]


class ResultSet(C.Structure):
    _fields_ = [("table", C.POINTER(Table)), ("cursor_blk", C.c_void_p), ("cursor_offset", C.c_size_t)]


class QueryOutput(C.Structure):
    _fields_ = [("status", C.c_int), ("results", ResultSet), ("error", C.c_char * 1024), ("n_rows_aff", C.c_size_t)]


class DatabaseStruct(C.Structure):
    _fields_ = [("tables", C.c_void_p), ("mutex", C.c_ubyte * 40)]


_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.database_open.argtypes = [C.POINTER(DatabaseStruct)]
        L.database_close.argtypes = [C.POINTER(DatabaseStruct)]
        L.database_close.restype = None
        L.query_execute.argtypes = [C.POINTER(DatabaseStruct), C.c_char_p]
        L.query_execute.restype = C.POINTER(QueryOutput)
        L.query_cur_step.argtypes = [C.POINTER(ResultSet)]
        L.query_column_int64.argtypes = [C.POINTER(ResultSet), C.c_int]
        L.query_column_int64.restype = C.c_int64
        L.query_column_double.argtypes = [C.POINTER(ResultSet), C.c_int]
        L.query_column_double.restype = C.c_double
        L.query_column_is_null.argtypes = [C.POINTER(ResultSet), C.c_int]
        L.query_column_is_null.restype = C.c_bool
        L.query_free.argtypes = [C.POINTER(QueryOutput)]
        L.query_free.restype = None
        L.midoridb_b200_last_path.argtypes = [C.POINTER(DatabaseStruct), C.POINTER(C.c_uint64)]
        L.midoridb_b200_scaffold_order.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def scaffold_order(keys):
    """keys (put order) -> keys in the reference's result-column order"""
    L = load_library()
    n = len(keys)
    arr = (C.c_char_p * n)(*[k.encode() for k in keys])
    pos = (C.c_int * n)()
    L.midoridb_b200_scaffold_order(arr, n, pos)
    return [k for _, k in sorted(zip(list(pos), keys))]


class QueryError(RuntimeError):
    pass


class Database:
    """`struct database` + the reference's call sequence (README.md:48-77)"""

    def __init__(self):
        self.L = load_library()
        self.db = DatabaseStruct()
        if self.L.database_open(C.byref(self.db)) != MIDORIDB_OK:
            raise RuntimeError("database_open failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if self.db is not None:
            self.L.database_close(C.byref(self.db))
            self.db = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def execute(self, sql):
        """CREATE / INSERT / UPDATE / DELETE; returns n_rows_aff"""
        out = self.L.query_execute(C.byref(self.db), sql.encode())
        st, msg, n = out.contents.status, out.contents.error.decode(errors="replace"), out.contents.n_rows_aff
        self.L.query_free(out)
        if st != ST_OK_EXECUTED:
            raise QueryError(msg.strip() or "statement failed")
        return n

    def query(self, sql):
        """SELECT through the cursor API; returns (column names, rows) with None for NULL"""
        out = self.L.query_execute(C.byref(self.db), sql.encode())
        try:
            if out.contents.status != ST_OK_WITH_RESULTS:
                raise QueryError(out.contents.error.decode(errors="replace").strip() or "query failed")
            rs = C.byref(out.contents.results)
            tbl = out.contents.results.table.contents
            ncols = tbl.column_count
            names = [tbl.columns[c].name.decode() for c in range(ncols)]
            types = [tbl.columns[c].type for c in range(ncols)]
            rows = []
            while self.L.query_cur_step(rs) == MIDORIDB_ROW:
                row = []
                for c in range(ncols):
                    if self.L.query_column_is_null(rs, c):
                        row.append(None)
                    elif types[c] == CT_DOUBLE:
                        row.append(self.L.query_column_double(rs, c))
                    else:
                        row.append(self.L.query_column_int64(rs, c))
                rows.append(tuple(row))
            return names, rows
        finally:
            self.L.query_free(out)

    def last_path(self):
        n = C.c_uint64()
        p = self.L.midoridb_b200_last_path(C.byref(self.db), C.byref(n))
        return p, n.value
