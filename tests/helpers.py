"""shared helpers for the parity tests (test infrastructure)"""
import json
import math
import os

import numpy as np

from midoridb_b200 import capi
from midoridb_b200.capi import CT_DOUBLE

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_select.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def table_arrays(tbl):
    """golden table dict -> (types, cells int64 [n, ncols], nulls uint8 [n, ncols], deleted uint8 [n])"""
    types = tbl["types"]
    n, ncols = len(tbl["rows"]), len(types)
    cells = np.zeros((n, ncols), dtype=np.int64)
    nulls = np.zeros((n, ncols), dtype=np.uint8)
    deleted = np.zeros(n, dtype=np.uint8)
    for r, row in enumerate(tbl["rows"]):
        deleted[r] = row["deleted"]
        for c, v in enumerate(row["v"]):
            if v is None:
                nulls[r, c] = 1
            elif types[c] == CT_DOUBLE:
                cells[r, c] = np.float64(v).view(np.int64)
            else:
                cells[r, c] = v
    return types, cells, nulls, deleted


def golden_pages(tbl):
    types, cells, nulls, deleted = table_arrays(tbl)
    return types, capi.pack_pages(types, cells, nulls, deleted)


def plan_from_dict(tables, d, flags=0):
    joins = []
    for j in d.get("joins", []):
        joins.append("cross" if j == "cross" else (tuple(j[0]), tuple(j[1])))
    return capi.make_plan(tables, joins=joins, pred=[tuple(p) for p in d.get("pred", [])],
                          group=[tuple(g) for g in d.get("group", [])], out=[tuple(o) for o in d["out"]], flags=flags)


def norm_row(row):
    return tuple(None if v is None else (float(v) if isinstance(v, float) else int(v)) for v in row)


def sort_key(row):
    return tuple((0, 0) if v is None else (1, v) for v in row)


def canon(rows):
    """row order canonicalised (north star: 'row order canonicalised before comparison')"""
    return sorted((norm_row(r) for r in rows), key=sort_key)


def rows_close(a, b, rel=1e-9):
    """bit-exact for ints / None, `rel` relative for floats (north star tolerance for DOUBLE SUM/AVG)"""
    if len(a) != len(b):
        return False
    for ra, rb in zip(a, b):
        if len(ra) != len(rb):
            return False
        for x, y in zip(ra, rb):
            if x is None or y is None:
                if x is not y:
                    return False
            elif isinstance(x, float) or isinstance(y, float):
                if not math.isclose(x, y, rel_tol=rel, abs_tol=0.0) and not (x == y):
                    return False
            elif x != y:
                return False
    return True


def canon_close(rows_a, rows_b, rel=1e-9):
    """compare two row sets after canonical ordering; floats may differ in the last bits, so sort on rounded keys"""
    def rkey(row):
        return tuple((0, 0) if v is None else (1, round(v, 6) if isinstance(v, float) else v) for v in row)
    a = sorted((norm_row(r) for r in rows_a), key=rkey)
    b = sorted((norm_row(r) for r in rows_b), key=rkey)
    return rows_close(a, b, rel)
