"""ctypes binding of include/midoridb_cuda.h (libmidoridb_cuda.so).

This is what a Python host would bind; the C host side (midoridb_b200/host) links the same symbols.
There is no fallback of any kind: if the shared library is missing or no CUDA device is present the
import of the library / `Backend()` raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmidoridb_cuda.so")

# keep in sync with include/midoridb_cuda.h
OK, EERROR, EINTERNAL, ENOMEM, EUNSUPPORTED, ECUDA = 0, -1, -2, -3, -16, -17
CT_VARCHAR, CT_INTEGER, CT_TINYINT, CT_DOUBLE, CT_DATE, CT_DATETIME = range(6)
PAGE_SIZE, ROW_HEADER = 4096, 24
MAX_TABLES, MAX_PRED, MAX_OUT, MAX_GROUP, MAX_HAVING, MAX_ORDER = 4, 64, 32, 2, 16, 4

P_COL, P_INT, P_DBL, P_NULL, P_CMP, P_AND, P_OR, P_XOR, P_ISNULL, P_ISNOTNULL, P_IN, P_NOTIN, P_BOOL, P_OUT = range(1, 15)
CMP_LT, CMP_GT, CMP_NE, CMP_EQ, CMP_LE, CMP_GE = 1, 2, 3, 4, 5, 6
OUT_COLUMN, OUT_COUNT_STAR, OUT_COUNT_COL, OUT_SUM, OUT_MIN, OUT_MAX, OUT_AVG = range(7)
PLAN_DISTRIBUTED, PLAN_NO_FASTPATH = 1, 2
GEN_UNIFORM_INT, GEN_UNIFORM_DBL, GEN_PERMUTATION, GEN_ZIPF, GEN_SEQUENCE = range(5)
PATH_GENERAL, PATH_SCAN_AGG, PATH_RADIX_JOINCOUNT, PATH_DIRECT_STAR, PATH_FUSED_MULTIWAY, PATH_DIRECT_COUNT = range(6)

EXPORTED_SYMBOLS = [
    "mdbcu_init", "mdbcu_shutdown", "mdbcu_last_error", "mdbcu_device_sync", "mdbcu_host_alloc", "mdbcu_host_free",
    "mdbcu_table_create", "mdbcu_table_drop", "mdbcu_table_append_pages", "mdbcu_table_append_page_ptrs",
    "mdbcu_table_reload_pages", "mdbcu_table_tombstone", "mdbcu_table_append_columns", "mdbcu_table_generate",
    "mdbcu_table_slots", "mdbcu_table_live_rows", "mdbcu_table_read_column", "mdbcu_table_column_device_ptr",
    "mdbcu_result_column_device_ptr",
    "mdbcu_select",
    "mdbcu_result_rows", "mdbcu_result_cols", "mdbcu_result_col_type", "mdbcu_result_fetch_columns",
    "mdbcu_result_page_count", "mdbcu_result_row_size", "mdbcu_result_fetch_pages", "mdbcu_result_free",
    "mdbcu_get_stats", "mdbcu_event_record", "mdbcu_event_elapsed_ms", "mdbcu_comm_unique_id", "mdbcu_comm_init", "mdbcu_comm_init_local", "mdbcu_comm_world", "mdbcu_table_sync_stats", "mdbcu_dist_describe", "mdbcu_dist_owner", "mdbcu_version",
]


class GenSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("null_permille", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64),
                ("param", C.c_double), ("seed", C.c_uint64)]


class PredOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("tbl", C.c_int32), ("col", C.c_int32),
                ("ival", C.c_int64), ("dval", C.c_double)]


class ColRef(C.Structure):
    _fields_ = [("tbl", C.c_int32), ("col", C.c_int32)]


class Join(C.Structure):
    _fields_ = [("left", ColRef), ("right", ColRef), ("cross", C.c_int32), ("_pad", C.c_int32)]


class Out(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ref", ColRef)]


class Order(C.Structure):
    _fields_ = [("out_col", C.c_int32), ("desc", C.c_int32)]


class Plan(C.Structure):
    _fields_ = [("n_tables", C.c_int32), ("tables", C.c_void_p * MAX_TABLES),
                ("n_joins", C.c_int32), ("joins", Join * (MAX_TABLES - 1)),
                ("n_pred", C.c_int32), ("pred", PredOp * MAX_PRED),
                ("n_group", C.c_int32), ("group", ColRef * MAX_GROUP),
                ("n_out", C.c_int32), ("out", Out * MAX_OUT),
                ("flags", C.c_uint32),
                ("distinct", C.c_int32), ("n_having", C.c_int32), ("having", PredOp * MAX_HAVING),
                ("n_order", C.c_int32), ("order", Order * MAX_ORDER),
                ("has_limit", C.c_int32), ("_pad2", C.c_int32), ("limit", C.c_int64), ("offset", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("phase_ms", C.c_double * 8), ("kernel_launches", C.c_uint64),
                ("total_kernel_launches", C.c_uint64), ("input_rows", C.c_uint64), ("result_rows", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("path", C.c_int32), ("_pad", C.c_int32),
                ("dominant_ms", C.c_double), ("dominant_bytes", C.c_uint64), ("exchange_bytes", C.c_uint64)]


class DistLayout(C.Structure):
    _fields_ = [("part_first", C.c_uint32 * 9), ("stream_cap", C.c_uint32), ("tail_cap", C.c_uint32), ("_pad", C.c_uint32),
                ("region_main_off", C.c_uint64), ("region_tail_off", C.c_uint64), ("region_cursor_off", C.c_uint64),
                ("region_bytes", C.c_uint64), ("arena_half_bytes", C.c_uint64)]


def dist_describe(nparts, world, global_rows, sms=148):
    """host-only: ownership and arena slot layout of a distributed radix join (mdbcu_dist_describe)"""
    L = load_library()
    out = DistLayout()
    rc = L.mdbcu_dist_describe(nparts, world, global_rows, sms, C.byref(out))
    if rc != OK:
        raise MdbError(rc, "mdbcu_dist_describe: bad arguments")
    return out


def dist_owner(partition, nparts, world):
    return load_library().mdbcu_dist_owner(partition, nparts, world)


class MdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("midoridb_cuda error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """dlopen libmidoridb_cuda.so and declare prototypes; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    L.mdbcu_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.mdbcu_shutdown.argtypes = [vp]
    L.mdbcu_shutdown.restype = None
    L.mdbcu_last_error.argtypes = [vp]
    L.mdbcu_last_error.restype = C.c_char_p
    L.mdbcu_device_sync.argtypes = [vp]
    L.mdbcu_host_alloc.argtypes = [vp, sz]
    L.mdbcu_host_alloc.restype = vp
    L.mdbcu_host_free.argtypes = [vp, vp]
    L.mdbcu_host_free.restype = None
    L.mdbcu_table_create.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(vp)]
    L.mdbcu_table_drop.argtypes = [vp]
    L.mdbcu_table_drop.restype = None
    L.mdbcu_table_append_pages.argtypes = [vp, vp, sz, sz]
    L.mdbcu_table_append_page_ptrs.argtypes = [vp, C.POINTER(vp), sz]
    L.mdbcu_table_reload_pages.argtypes = [vp, sz, C.POINTER(vp), sz]
    L.mdbcu_table_tombstone.argtypes = [vp, vp, vp, sz]
    L.mdbcu_table_append_columns.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(vp)]
    L.mdbcu_table_generate.argtypes = [vp, u64, u64, C.POINTER(GenSpec)]
    L.mdbcu_table_slots.argtypes = [vp]
    L.mdbcu_table_slots.restype = u64
    L.mdbcu_table_live_rows.argtypes = [vp]
    L.mdbcu_table_live_rows.restype = u64
    L.mdbcu_table_read_column.argtypes = [vp, C.c_int, u64, u64, vp, vp]
    L.mdbcu_table_column_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(u64)]
    L.mdbcu_result_column_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.mdbcu_select.argtypes = [vp, C.POINTER(Plan), C.POINTER(vp)]
    L.mdbcu_result_rows.argtypes = [vp]
    L.mdbcu_result_rows.restype = u64
    L.mdbcu_result_cols.argtypes = [vp]
    L.mdbcu_result_col_type.argtypes = [vp, C.c_int]
    L.mdbcu_result_fetch_columns.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mdbcu_result_page_count.argtypes = [vp]
    L.mdbcu_result_page_count.restype = sz
    L.mdbcu_result_row_size.argtypes = [vp]
    L.mdbcu_result_row_size.restype = sz
    L.mdbcu_result_fetch_pages.argtypes = [vp, vp, sz]
    L.mdbcu_result_free.argtypes = [vp]
    L.mdbcu_result_free.restype = None
    L.mdbcu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mdbcu_event_record.argtypes = [vp, C.c_int]
    L.mdbcu_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.mdbcu_comm_unique_id.argtypes = [vp, vp]
    L.mdbcu_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.mdbcu_comm_init_local.argtypes = [C.POINTER(vp), C.c_int]
    L.mdbcu_comm_world.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mdbcu_dist_describe.argtypes = [C.c_int, C.c_int, u64, C.c_int, C.POINTER(DistLayout)]
    L.mdbcu_dist_owner.argtypes = [C.c_uint32, C.c_int, C.c_int]
    L.mdbcu_table_sync_stats.argtypes = [vp]
    L.mdbcu_version.restype = C.c_char_p
    _lib = L
    return L


def row_size_of(col_types):
    """table_calc_row_size (src/primitive/row.c:21): 24-byte header + column widths"""
    return ROW_HEADER + sum(1 if t == CT_TINYINT else 8 for t in col_types)


def pack_pages(col_types, cells, nulls=None, deleted=None):
    """Lay rows out exactly as the reference's table_insert_row does (src/primitive/row.c:26-98):
    returns a contiguous uint8 array of n_pages x 4096 page data areas.  Host-side helper used by the
    tests and the e2e bench leg to produce reference-format inputs without the reference."""
    ncols = len(col_types)
    cells = np.ascontiguousarray(cells)
    if cells.ndim == 1:
        cells = cells.reshape(-1, 1)
    n = cells.shape[0]
    assert cells.shape[1] == ncols
    rs = row_size_of(col_types)
    rpp = (PAGE_SIZE - 1) // rs
    slots = PAGE_SIZE // rs
    n_pages = max(1, (n + rpp - 1) // rpp) if n else 0
    pages = np.zeros((n_pages, PAGE_SIZE), dtype=np.uint8)
    if n_pages == 0:
        return pages
    # every slot starts out empty (table_datablock_init, src/primitive/table.c:124-132)
    slot_view = pages[:, :slots * rs].reshape(n_pages, slots, rs)
    slot_view[:, :, 0] = 1
    page_of = np.arange(n) // rpp
    slot_of = np.arange(n) % rpp
    rows = np.zeros((n, rs), dtype=np.uint8)
    if deleted is not None:
        rows[:, 1] = np.asarray(deleted, dtype=np.uint8)
    off = ROW_HEADER
    raw = cells.view(np.int64) if cells.dtype == np.float64 else cells.astype(np.int64, copy=False)
    for c, t in enumerate(col_types):
        w = 1 if t == CT_TINYINT else 8
        col_bytes = np.ascontiguousarray(raw[:, c]).view(np.uint8).reshape(n, 8)[:, :w]
        if nulls is not None:
            isnull = np.asarray(nulls)[:, c].astype(bool)
            rows[:, 2 + c // 8] |= (isnull.astype(np.uint8) << (c % 8))
            col_bytes = np.where(isnull[:, None], 0, col_bytes)
        rows[:, off:off + w] = col_bytes
        off += w
    slot_view[page_of, slot_of, :] = rows
    return pages


def unpack_pages(pages, ncols):
    """decode result page images (24-byte header + 8-byte cells) with the executor idiom; returns (cells, nulls)"""
    pages = np.asarray(pages, dtype=np.uint8).reshape(-1, PAGE_SIZE)
    rs = ROW_HEADER + 8 * ncols
    slots = PAGE_SIZE // rs
    out_cells, out_nulls = [], []
    for p in range(pages.shape[0]):
        v = pages[p, :slots * rs].reshape(slots, rs)
        for s in range(slots):
            if v[s, 0]:
                break
            if v[s, 1]:
                continue
            out_cells.append(v[s, ROW_HEADER:].copy().view(np.int64))
            out_nulls.append([(v[s, 2 + c // 8] >> (c % 8)) & 1 for c in range(ncols)])
    cells = np.array(out_cells, dtype=np.int64).reshape(-1, ncols)
    nulls = np.array(out_nulls, dtype=np.uint8).reshape(-1, ncols)
    return cells, nulls


class Table:
    def __init__(self, backend, name, col_types, handle):
        self.backend, self.name, self.col_types, self.handle = backend, name, list(col_types), handle

    def _check(self, rc):
        self.backend._check(rc)

    def append_pages(self, pages, stride=PAGE_SIZE):
        pages = np.ascontiguousarray(pages, dtype=np.uint8)
        n_pages = pages.size // stride
        self._check(self.backend.L.mdbcu_table_append_pages(self.handle, pages.ctypes.data, n_pages, stride))

    def append_page_ptrs(self, ptrs):
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(self.backend.L.mdbcu_table_append_page_ptrs(self.handle, arr, len(ptrs)))

    def reload_pages(self, first_page, ptrs):
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(self.backend.L.mdbcu_table_reload_pages(self.handle, first_page, arr, len(ptrs)))

    def tombstone(self, page_idx, slot_idx):
        p = np.ascontiguousarray(page_idx, dtype=np.uint64)
        s = np.ascontiguousarray(slot_idx, dtype=np.uint32)
        self._check(self.backend.L.mdbcu_table_tombstone(self.handle, p.ctypes.data, s.ctypes.data, p.size))

    def append_columns(self, columns, nulls=None):
        cols = [np.ascontiguousarray(c) for c in columns]
        for c in cols:
            assert c.dtype in (np.int64, np.float64) and c.ndim == 1
        n = cols[0].size
        data = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        nl = None
        keep = []
        if nulls is not None:
            keep = [None if x is None else np.ascontiguousarray(x, dtype=np.uint8) for x in nulls]
            nl = (C.c_void_p * len(cols))(*[None if x is None else x.ctypes.data for x in keep])
        self._check(self.backend.L.mdbcu_table_append_columns(self.handle, n, data, nl))

    def generate(self, n_rows, specs, row_offset=0):
        arr = (GenSpec * len(specs))(*specs)
        self._check(self.backend.L.mdbcu_table_generate(self.handle, n_rows, row_offset, arr))

    def sync_stats(self):
        """collective: make the column statistics global across the ranks' shards"""
        self._check(self.backend.L.mdbcu_table_sync_stats(self.handle))

    @property
    def slots(self):
        return self.backend.L.mdbcu_table_slots(self.handle)

    def live_rows(self):
        return self.backend.L.mdbcu_table_live_rows(self.handle)

    def read_column(self, col, first=0, n=None):
        n = self.slots - first if n is None else n
        dt = np.float64 if self.col_types[col] == CT_DOUBLE else np.int64
        cells = np.zeros(n, dtype=dt)
        valid = np.zeros(n, dtype=np.uint8)
        self._check(self.backend.L.mdbcu_table_read_column(self.handle, col, first, n, cells.ctypes.data, valid.ctypes.data))
        return cells, valid

    def device_ptr(self, col):
        """(device address, slots) of one mirrored column: zero-copy hand-over to a harness with its own CUDA code"""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.backend.L.mdbcu_table_column_device_ptr(self.handle, col, C.byref(ptr), C.byref(n)))
        return ptr.value or 0, n.value

    def drop(self):
        if self.handle:
            self.backend.L.mdbcu_table_drop(self.handle)
            self.handle = None


class DeviceArray:
    """__cuda_array_interface__ view of library-owned device memory (torch.as_tensor(DeviceArray(...), device="cuda"))"""

    def __init__(self, ptr, n, typestr="<i8", owner=None):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class Result:
    def __init__(self, backend, handle):
        self.backend, self.handle = backend, handle
        L = backend.L
        self.nrows = L.mdbcu_result_rows(handle)
        self.ncols = L.mdbcu_result_cols(handle)
        self.types = [L.mdbcu_result_col_type(handle, c) for c in range(self.ncols)]

    def fetch_columns(self):
        """returns (cells[list of arrays, int64 or float64 per column type], nulls[list of uint8 arrays])"""
        cells = [np.zeros(self.nrows, dtype=np.float64 if t == CT_DOUBLE else np.int64) for t in self.types]
        nulls = [np.zeros(self.nrows, dtype=np.uint8) for _ in self.types]
        cp = (C.c_void_p * max(self.ncols, 1))(*[c.ctypes.data for c in cells])
        npn = (C.c_void_p * max(self.ncols, 1))(*[n.ctypes.data for n in nulls])
        self.backend._check(self.backend.L.mdbcu_result_fetch_columns(self.handle, cp, npn))
        return cells, nulls

    def fetch_columns_into(self, cell_ptrs, null_ptrs=None):
        """copy result columns into caller-owned buffers (raw addresses, e.g. page-locked memory)"""
        cp = (C.c_void_p * max(self.ncols, 1))(*cell_ptrs)
        npn = None if null_ptrs is None else (C.c_void_p * max(self.ncols, 1))(*null_ptrs)
        self.backend._check(self.backend.L.mdbcu_result_fetch_columns(self.handle, cp, npn))

    def device_ptr(self, col):
        """device address of one result column (rows in device order)"""
        ptr, nul = C.c_void_p(), C.c_void_p()
        self.backend._check(self.backend.L.mdbcu_result_column_device_ptr(self.handle, col, C.byref(ptr), C.byref(nul)))
        return ptr.value or 0

    def page_count(self):
        return self.backend.L.mdbcu_result_page_count(self.handle)

    def fetch_pages(self, out=None):
        L = self.backend.L
        n = L.mdbcu_result_page_count(self.handle)
        if out is None:
            pages = np.zeros((n, PAGE_SIZE), dtype=np.uint8)
        else:
            assert out.size >= n * PAGE_SIZE
            pages = out.reshape(-1)[:n * PAGE_SIZE].reshape(n, PAGE_SIZE)
        self.backend._check(L.mdbcu_result_fetch_pages(self.handle, pages.ctypes.data, n))
        return pages

    def rows(self):
        """rows as tuples (None for NULL), doubles as Python floats"""
        cells, nulls = self.fetch_columns()
        out = []
        for r in range(self.nrows):
            out.append(tuple(None if nulls[c][r] else cells[c][r].item() for c in range(self.ncols)))
        return out

    def free(self):
        if self.handle:
            self.backend.L.mdbcu_result_free(self.handle)
            self.handle = None


class Backend:
    """one CUDA context of libmidoridb_cuda.so (one per process / GPU)"""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.mdbcu_init(device, C.byref(h))
        if rc != OK:
            raise MdbError(rc, (self.L.mdbcu_last_error(None) or b"").decode())
        self.ctx = h
        self.device = device

    def _check(self, rc):
        if rc != OK:
            raise MdbError(rc, (self.L.mdbcu_last_error(self.ctx) or b"").decode())

    def close(self):
        if self.ctx:
            self.L.mdbcu_shutdown(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._check(self.L.mdbcu_device_sync(self.ctx))

    def create_table(self, name, col_types):
        types = (C.c_int32 * len(col_types))(*col_types)
        h = C.c_void_p()
        self._check(self.L.mdbcu_table_create(self.ctx, name.encode(), len(col_types), types, C.byref(h)))
        return Table(self, name, col_types, h)

    def select(self, plan):
        h = C.c_void_p()
        self._check(self.L.mdbcu_select(self.ctx, C.byref(plan), C.byref(h)))
        return Result(self, h)

    def stats(self):
        s = Stats()
        self._check(self.L.mdbcu_get_stats(self.ctx, C.byref(s)))
        return s

    def host_array(self, nbytes):
        """uint8 numpy array over page-locked host memory (freed with host_free(arr))"""
        p = self.L.mdbcu_host_alloc(self.ctx, nbytes)
        if not p:
            raise MdbError(ENOMEM, (self.L.mdbcu_last_error(self.ctx) or b"").decode())
        arr = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p
        return arr

    def host_free(self, arr):
        p = getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)
        if p:
            self.L.mdbcu_host_free(self.ctx, p)

    def event_record(self, slot):
        self._check(self.L.mdbcu_event_record(self.ctx, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        self._check(self.L.mdbcu_event_elapsed_ms(self.ctx, a, b, C.byref(ms)))
        return ms.value

    @staticmethod
    def _prefer_bundled_nccl():
        """a process that imports torch later must find the NCCL build torch was linked against, not the system one: load
        the wheel's copy first (and tell the library where it is) when it exists"""
        if os.environ.get("MDBCU_NCCL_LIB"):
            return
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                os.environ["MDBCU_NCCL_LIB"] = path
                try:
                    C.CDLL(path, mode=C.RTLD_GLOBAL)
                except OSError:
                    del os.environ["MDBCU_NCCL_LIB"]
                return

    def comm_unique_id(self):
        self._prefer_bundled_nccl()
        buf = (C.c_ubyte * 128)()
        self._check(self.L.mdbcu_comm_unique_id(self.ctx, buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        self._prefer_bundled_nccl()
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self.L.mdbcu_comm_init(self.ctx, rank, world, buf))


def comm_init_local(backends):
    """make the given Backends (same process) ranks 0..n-1 of one in-process communicator (loop-back mode when they share a GPU);
    afterwards every collective must be called by all of them concurrently, one host thread each"""
    arr = (C.c_void_p * len(backends))(*[b.ctx for b in backends])
    backends[0]._check(backends[0].L.mdbcu_comm_init_local(arr, len(backends)))


def make_plan(tables, joins=(), pred=(), group=(), out=(), flags=0, distinct=False, having=(), order=(), limit=None, offset=0):
    """Build a `struct mdbcu_plan`.
    having: postfix program like `pred` with ("out", i) = result column i;  order: list of (result column, descending);
    limit / offset: LIMIT [offset,] count
    tables: list of handles (Table objects, raw pointers or oracle tables exposing `.handle`)
    joins:  list of ((ltbl, lcol), (rtbl, rcol)) or "cross"
    pred:   postfix list of tuples: ("col", tbl, col) ("int", v) ("dbl", v) ("null",) ("bool", v)
            ("cmp", k) ("and",) ("or",) ("xor",) ("isnull",) ("isnotnull",) ("in", n) ("notin", n)
    group:  list of (tbl, col);  out: list of (kind, tbl, col) or (OUT_COUNT_STAR,)"""
    p = Plan()
    p.n_tables = len(tables)
    for i, t in enumerate(tables):
        h = getattr(t, "handle", t)
        p.tables[i] = h.value if isinstance(h, C.c_void_p) else h
    p.n_joins = len(joins)
    for i, j in enumerate(joins):
        if j == "cross":
            p.joins[i].cross = 1
        else:
            (lt, lc), (rt, rc) = j
            p.joins[i].left = ColRef(lt, lc)
            p.joins[i].right = ColRef(rt, rc)
    names = {"col": P_COL, "int": P_INT, "dbl": P_DBL, "null": P_NULL, "bool": P_BOOL, "cmp": P_CMP, "and": P_AND,
             "or": P_OR, "xor": P_XOR, "isnull": P_ISNULL, "isnotnull": P_ISNOTNULL, "in": P_IN, "notin": P_NOTIN, "out": P_OUT}

    def fill(o, op):
        o.op = names[op[0]]
        if op[0] == "col":
            o.tbl, o.col = op[1], op[2]
        elif op[0] == "out":
            o.col = int(op[1])
        elif op[0] in ("int", "bool"):
            o.ival = int(op[1])
        elif op[0] == "dbl":
            o.dval = float(op[1])
        elif op[0] in ("cmp", "in", "notin"):
            o.arg = int(op[1])

    p.n_pred = len(pred)
    for i, op in enumerate(pred):
        fill(p.pred[i], op)
    p.distinct = 1 if distinct else 0
    p.n_having = len(having)
    for i, op in enumerate(having):
        fill(p.having[i], op)
    p.n_order = len(order)
    for i, (c, desc) in enumerate(order):
        p.order[i] = Order(int(c), 1 if desc else 0)
    if limit is not None:
        p.has_limit, p.limit, p.offset = 1, int(limit), int(offset)
    p.n_group = len(group)
    for i, (t, c) in enumerate(group):
        p.group[i] = ColRef(t, c)
    p.n_out = len(out)
    for i, o in enumerate(out):
        p.out[i].kind = o[0]
        if len(o) > 1:
            p.out[i].ref = ColRef(o[1], o[2])
    p.flags = flags
    return p
