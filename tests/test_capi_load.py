"""CPU tests: the C-ABI library loads, exports every symbol include/midoridb_cuda.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from midoridb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "midoridb_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(mdbcu_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    L = capi.load_library()
    for name in declared:
        assert hasattr(L, name), "libmidoridb_cuda.so does not export %s" % name


def test_plan_struct_layout_matches_header():
    # sizes the C side relies on (struct mdbcu_pred_op = 32 bytes, mdbcu_gen_spec = 40 bytes, ...)
    assert C.sizeof(capi.PredOp) == 32
    assert C.sizeof(capi.GenSpec) == 40
    assert C.sizeof(capi.ColRef) == 8
    assert C.sizeof(capi.Join) == 24
    assert C.sizeof(capi.Out) == 12
    assert C.sizeof(capi.Stats) == 8 + 64 + 8 * 5 + 8 + 8 + 8 + 8


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_init_fails_loudly_without_gpu():
    L = capi.load_library()
    h = C.c_void_p()
    rc = L.mdbcu_init(0, C.byref(h))
    assert rc == capi.ECUDA and not h.value
    msg = L.mdbcu_last_error(None).decode()
    assert "no CPU fallback" in msg
    with pytest.raises(capi.MdbError):
        capi.Backend(0)


def test_pack_unpack_roundtrip():
    import numpy as np
    rng = np.random.default_rng(1)
    cells = rng.integers(-10**12, 10**12, (500, 3), dtype=np.int64)
    nulls = (rng.random((500, 3)) < 0.3).astype(np.uint8)
    pages = capi.pack_pages([capi.CT_INTEGER] * 3, cells, nulls)
    c2, n2 = capi.unpack_pages(pages, 3)
    assert np.array_equal(n2, nulls)
    assert np.array_equal(np.where(nulls, 0, cells), c2)
