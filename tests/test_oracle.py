"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors, against the reference itself
(oracle/_ref, when present) and against sqlite3 for the extensions the reference cannot run."""
import ctypes
import os
import sqlite3

import numpy as np
import pytest

from midoridb_b200 import capi
from midoridb_b200.capi import (CT_DOUBLE, CT_INTEGER, OUT_AVG, OUT_COLUMN, OUT_COUNT_COL, OUT_COUNT_STAR, OUT_MAX, OUT_MIN,
                                OUT_SUM)
from oracle import oracle, refdb
from tests import helpers

CASES = helpers.load_golden()
needs_ref = pytest.mark.skipif(not refdb.available(), reason="oracle/_ref not built (reference tree absent)")


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(case):
    """oracle == reference's known answers, in the reference's row order"""
    tables = []
    for tbl in case["tables"]:
        types, pages = helpers.golden_pages(tbl)
        t = oracle.OracleTable(types)
        t.append_pages(pages)
        tables.append(t)
    plan = helpers.plan_from_dict(tables, case["plan"])
    _, cells, nulls = oracle.select(plan)
    got = [helpers.norm_row(r) for r in oracle.rows_of(cells, nulls)]
    want = [helpers.norm_row(r) for r in case["rows"]]
    assert got == want
    for t in tables:
        t.free()


def test_golden_covers_reference_tests():
    names = {c["name"] for c in CASES}
    for i in range(1, 13):
        assert "test_select_%d" % i in names


@needs_ref
def test_pack_pages_matches_reference_pages():
    """storage layout pin (tests/primitive/row.c:15-136): our packer == bytes written by table_insert_row"""
    rng = np.random.default_rng(7)
    for ncols, n in ((1, 300), (2, 260), (5, 77), (13, 40)):
        types = [CT_INTEGER if c % 2 == 0 else CT_DOUBLE for c in range(ncols)]
        cells = rng.integers(-2**40, 2**40, (n, ncols), dtype=np.int64)
        nulls = (rng.random((n, ncols)) < 0.2).astype(np.uint8)
        with refdb.RefDatabase() as db:
            t = db.create_table("T", ["c%d" % c for c in range(ncols)], types)
            db.append(t, cells, nulls)
            ptrs, rs = db.page_ptrs(t)
            assert rs == capi.row_size_of(types)
            ref_pages = np.stack([np.frombuffer((ctypes.c_ubyte * 4096).from_address(p), dtype=np.uint8) for p in ptrs])
            mine = capi.pack_pages(types, cells, nulls)
            assert mine.shape == ref_pages.shape
            rpp = 4095 // rs
            # compare every slot the executor can reach (the bytes behind the last slot are never initialised
            # by the reference: datablock_alloc mallocs, src/primitive/datablock.c:17)
            used = (4096 // rs) * rs
            assert np.array_equal(mine[:, :used], ref_pages[:, :used])
            assert len(ptrs) == (n + rpp - 1) // rpp


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_reference_random(seed):
    """randomised inputs inside the reference's correct domain: one-page join outputs with
    duplicates and NULLs, and multi-page unique-key joins"""
    rng = np.random.default_rng(100 + seed)
    if seed % 2 == 0:
        na, nb, dom = int(rng.integers(3, 12)), int(rng.integers(5, 30)), int(rng.integers(3, 15))
        a = rng.integers(0, dom, na)
        b = rng.integers(0, dom, nb)
    else:
        n = int(rng.integers(150, 500))
        a = rng.permutation(n)
        b = rng.permutation(n)[: int(n * 0.7)]
        na, nb = a.size, b.size
    an = (rng.random(na) < 0.1).astype(np.uint8)
    bn = (rng.random(nb) < 0.1).astype(np.uint8)
    sql = "SELECT id_a, COUNT(*) FROM A INNER JOIN B ON A.id_a = B.id_b GROUP BY id_a;"
    with refdb.RefDatabase() as db:
        ta = db.create_table("A", ["id_a"], [CT_INTEGER])
        tb = db.create_table("B", ["id_b"], [CT_INTEGER])
        db.append(ta, a.astype(np.int64), an)
        db.append(tb, b.astype(np.int64), bn)
        ref = db.query(sql)
        ref_rows = [(int(ref.cells[r, 0]), int(ref.cells[r, 1])) for r in range(ref.cells.shape[0])]
    oa, ob = oracle.OracleTable([CT_INTEGER]), oracle.OracleTable([CT_INTEGER])
    oa.append_pages(capi.pack_pages([CT_INTEGER], a.astype(np.int64), an.reshape(-1, 1)))
    ob.append_pages(capi.pack_pages([CT_INTEGER], b.astype(np.int64), bn.reshape(-1, 1)))
    plan = capi.make_plan([oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    _, cells, nulls = oracle.select(plan)
    got = oracle.rows_of(cells, nulls)
    if len(ref_rows) <= 85:
        assert got == ref_rows  # same order as the reference
    else:
        assert helpers.canon(got) == helpers.canon(ref_rows)


def _sqlite_rows(tables, sql):
    con = sqlite3.connect(":memory:")
    for name, cols, rows in tables:
        con.execute("CREATE TABLE %s (%s)" % (name, ", ".join(cols)))
        con.executemany("INSERT INTO %s VALUES (%s)" % (name, ",".join("?" * len(cols))), rows)
    return [tuple(r) for r in con.execute(sql).fetchall()]


def _oracle_table(types, rows):
    t = oracle.OracleTable(types)
    n, nc = len(rows), len(types)
    cols, nulls = [], []
    for c in range(nc):
        isn = np.array([r[c] is None for r in rows], dtype=np.uint8)
        if types[c] == CT_DOUBLE:
            col = np.array([0.0 if r[c] is None else r[c] for r in rows], dtype=np.float64)
        else:
            col = np.array([0 if r[c] is None else r[c] for r in rows], dtype=np.int64)
        cols.append(col)
        nulls.append(isn)
    t.append_columns(cols, nulls)
    return t


def test_oracle_extensions_match_sqlite():
    """parity UNPINNED by the reference (no SUM/MIN/MAX/AVG, D2, D3): oracle follows SQL, checked against sqlite3"""
    rng = np.random.default_rng(5)
    n = 3000
    a_rows = [(int(k), None if rng.random() < 0.1 else float(np.round(rng.random() * 100, 3))) for k in rng.integers(0, 200, n)]
    b_rows = [(int(k), None if rng.random() < 0.1 else int(rng.integers(-1000, 1000))) for k in rng.integers(0, 200, 500)]
    c_rows = [(int(k), int(rng.integers(0, 50))) for k in rng.permutation(200)[:150]]
    I, D = CT_INTEGER, CT_DOUBLE
    ta, tb, tc = _oracle_table([I, D], a_rows), _oracle_table([I, I], b_rows), _oracle_table([I, I], c_rows)
    sq = [("A", ["id", "x"], a_rows), ("B", ["id", "y"], b_rows), ("C", ["id", "z"], c_rows)]

    # multi-page GROUP BY with every aggregate (the reference loses counts here, D2)
    plan = capi.make_plan([ta], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_COUNT_COL, 0, 1),
                                                    (OUT_SUM, 0, 1), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1), (OUT_AVG, 0, 1)])
    _, cells, nulls = oracle.select(plan)
    want = _sqlite_rows(sq, "SELECT id, COUNT(*), COUNT(x), SUM(x), MIN(x), MAX(x), AVG(x) FROM A GROUP BY id")
    assert helpers.canon_close(oracle.rows_of(cells, nulls), want)

    # three-way join + WHERE + GROUP BY SUM/AVG (config 4 shape; the reference's 3-way join is broken, D3)
    plan = capi.make_plan([ta, tb, tc], joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
                          pred=[("col", 0, 1), ("dbl", 25.0), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
                          group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_SUM, 0, 1), (OUT_AVG, 2, 1), (OUT_COUNT_STAR,)])
    _, cells, nulls = oracle.select(plan)
    want = _sqlite_rows(sq, "SELECT A.id, SUM(A.x), AVG(C.z), COUNT(*) FROM A JOIN B ON A.id = B.id JOIN C ON A.id = C.id "
                            "WHERE A.x >= 25.0 AND B.y < 500 GROUP BY A.id")
    assert helpers.canon_close(oracle.rows_of(cells, nulls), want)

    # filter + aggregate scan without GROUP BY (config 2 shape), integer SUM/MIN/MAX
    plan = capi.make_plan([tb], pred=[("col", 0, 0), ("int", 50), ("cmp", 6), ("col", 0, 0), ("int", 149), ("cmp", 5), ("and",)],
                          out=[(OUT_COUNT_STAR,), (OUT_SUM, 0, 1), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1), (OUT_AVG, 0, 1)])
    _, cells, nulls = oracle.select(plan)
    want = _sqlite_rows(sq, "SELECT COUNT(*), SUM(y), MIN(y), MAX(y), AVG(y) FROM B WHERE id >= 50 AND id <= 149")
    assert helpers.canon_close(oracle.rows_of(cells, nulls), want)

    # IN = any-of (SQL; the reference evaluates it as all-of, D4)
    plan = capi.make_plan([tb], pred=[("col", 0, 0), ("int", 3), ("int", 5), ("int", 8), ("in", 3)], out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 0, 1)])
    _, cells, nulls = oracle.select(plan)
    want = _sqlite_rows(sq, "SELECT id, y FROM B WHERE id IN (3, 5, 8)")
    assert helpers.canon_close(oracle.rows_of(cells, nulls), want)


def test_join_count_port_matches_select():
    rng = np.random.default_rng(9)
    a = rng.integers(0, 5000, 20000).astype(np.int64)
    b = rng.integers(0, 5000, 30000).astype(np.int64)
    keys, cnts = oracle.join_count_groups(a, b, want_output=True)
    ta, tb = oracle.OracleTable([CT_INTEGER]), oracle.OracleTable([CT_INTEGER])
    ta.append_columns([a])
    tb.append_columns([b])
    plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    _, cells, _ = oracle.select(plan)
    assert sorted(zip(keys.tolist(), cnts.tolist())) == sorted(zip(cells[0].tolist(), cells[1].tolist()))
    assert int(cnts.sum()) == int(np.sum(np.bincount(a, minlength=5000) * np.bincount(b, minlength=5000)))


TAIL_CASES = [
    # (name, plan keyword arguments on table A(id, x) [+ B(id, y)], equivalent sqlite query, ordered comparison?)
    ("order_by_count_desc_key", dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)], order=[(1, True), (0, False)]),
     "SELECT id, COUNT(*) FROM A GROUP BY id ORDER BY COUNT(*) DESC, id", True),
    ("order_limit_offset", dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_SUM, 0, 1)], order=[(0, True)], limit=7, offset=5),
     "SELECT id, SUM(x) FROM A GROUP BY id ORDER BY id DESC LIMIT 7 OFFSET 5", True),
    ("having_count_and_sum", dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_SUM, 0, 1)],
                                  having=[("out", 1), ("int", 12), ("cmp", 2), ("out", 2), ("dbl", 900.0), ("cmp", 1), ("and",)]),
     "SELECT id, COUNT(*), SUM(x) FROM A GROUP BY id HAVING COUNT(*) > 12 AND SUM(x) < 900.0", False),
    ("having_null_aggregate", dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_MIN, 0, 1)], having=[("out", 1), ("isnull",)]),
     "SELECT id, MIN(x) FROM A GROUP BY id HAVING MIN(x) IS NULL", False),
    ("distinct_projection", dict(out=[(OUT_COLUMN, 0, 0)], distinct=True), "SELECT DISTINCT id FROM A", False),
    ("distinct_two_columns_with_nulls", dict(out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 0, 1)], distinct=True, order=[(0, False), (1, False)]),
     "SELECT DISTINCT id, x FROM A ORDER BY id, x", True),
    ("order_by_nullable_double_asc", dict(out=[(OUT_COLUMN, 0, 1), (OUT_COLUMN, 0, 0)], order=[(0, False), (1, False)], limit=40),
     "SELECT x, id FROM A ORDER BY x, id LIMIT 40", True),
    ("order_by_nullable_double_desc", dict(out=[(OUT_COLUMN, 0, 1), (OUT_COLUMN, 0, 0)], order=[(0, True), (1, True)], limit=40, offset=2950),
     "SELECT x, id FROM A ORDER BY x DESC, id DESC LIMIT 40 OFFSET 2950", True),
    ("join_group_having_order_limit", dict(joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_MAX, 1, 1)],
                                           having=[("out", 1), ("int", 30), ("cmp", 6)], order=[(1, True), (0, False)], limit=10),
     "SELECT A.id, COUNT(*), MAX(B.y) FROM A JOIN B ON A.id = B.id GROUP BY A.id HAVING COUNT(*) >= 30 ORDER BY COUNT(*) DESC, A.id LIMIT 10", True),
    ("limit_zero", dict(out=[(OUT_COLUMN, 0, 0)], order=[(0, False)], limit=0), "SELECT id FROM A ORDER BY id LIMIT 0", True),
    ("where_distinct_order", dict(pred=[("col", 0, 0), ("int", 100), ("cmp", 1)], out=[(OUT_COLUMN, 0, 0)], distinct=True, order=[(0, True)]),
     "SELECT DISTINCT id FROM A WHERE id < 100 ORDER BY id DESC", True),
]


def tail_tables():
    rng = np.random.default_rng(17)
    a_rows = [(int(k), None if rng.random() < 0.1 else float(np.round(rng.random() * 100 - 30, 1))) for k in rng.integers(0, 200, 3000)]
    a_rows[5] = (a_rows[5][0], -0.0)
    a_rows[6] = (a_rows[5][0], 0.0)  # -0.0 and 0.0 are one value for DISTINCT and ORDER BY
    b_rows = [(int(k), None if rng.random() < 0.1 else int(rng.integers(-1000, 1000))) for k in rng.integers(0, 200, 500)]
    return a_rows, b_rows


@pytest.mark.parametrize("case", TAIL_CASES, ids=[c[0] for c in TAIL_CASES])
def test_oracle_tail_operators_match_sqlite(case):
    """HAVING / DISTINCT / ORDER BY / LIMIT: the reference accepts and ignores them (midorisql.y:180-196,203; executor_select.c:1723),
    so parity is UNPINNED by the reference; the oracle's restatement of SQL semantics is checked against sqlite3 itself"""
    name, kw, sql, ordered = case
    a_rows, b_rows = tail_tables()
    ta, tb = _oracle_table([CT_INTEGER, CT_DOUBLE], a_rows), _oracle_table([CT_INTEGER, CT_INTEGER], b_rows)
    tables = [ta, tb] if kw.get("joins") else [ta]
    _, cells, nulls = oracle.select(capi.make_plan(tables, **kw))
    got = oracle.rows_of(cells, nulls)
    want = _sqlite_rows([("A", ["id", "x"], a_rows), ("B", ["id", "y"], b_rows)], sql)
    if ordered:
        assert helpers.rows_close([helpers.norm_row(r) for r in got], [helpers.norm_row(r) for r in want])
    else:
        assert helpers.canon_close(got, want)
