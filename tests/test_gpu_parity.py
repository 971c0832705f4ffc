"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (libmidoridb_cuda.so), against
  * the reference's known answers (tests/golden/reference_select.json),
  * the CPU oracle on seeded random inputs (bit-exact for ints, 1e-9 relative for DOUBLE SUM/AVG),
  * size-independent properties at larger sizes.
Nothing here reads /root/reference."""
import numpy as np
import pytest

from midoridb_b200 import capi
from midoridb_b200.capi import (CT_DOUBLE, CT_INTEGER, CT_TINYINT, OUT_AVG, OUT_COLUMN, OUT_COUNT_COL, OUT_COUNT_STAR, OUT_MAX,
                                OUT_MIN, OUT_SUM, PLAN_NO_FASTPATH)
from oracle import oracle
from tests import helpers

pytestmark = pytest.mark.gpu
CASES = helpers.load_golden()
I, D = CT_INTEGER, CT_DOUBLE


@pytest.fixture(scope="module")
def be():
    b = capi.Backend(0)
    yield b
    b.close()


def both_tables(be, types, columns, nulls=None, paged=False):
    """the same data in a device table and an oracle table"""
    g = be.create_table("t", types)
    o = oracle.OracleTable(types)
    if paged:
        cells = np.stack([c.view(np.int64) if c.dtype == np.float64 else c for c in columns], axis=1)
        nl = None if nulls is None else np.stack([np.zeros(len(columns[0]), np.uint8) if x is None else x for x in nulls], axis=1)
        pages = capi.pack_pages(types, cells, nl)
        g.append_pages(pages)
        o.append_pages(pages)
    else:
        g.append_columns(columns, nulls)
        o.append_columns(columns, nulls)
    return g, o


def run_both(be, gt, ot, flags=0, **kw):
    gp = capi.make_plan(gt, flags=flags, **kw)
    op = capi.make_plan(ot, **kw)
    res = be.select(gp)
    grows = res.rows()
    pages = res.fetch_pages()
    res.free()
    _, cells, nulls = oracle.select(op)
    return grows, oracle.rows_of(cells, nulls), pages, be.stats()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_golden(be, case):
    tables = []
    for tbl in case["tables"]:
        types, pages = helpers.golden_pages(tbl)
        t = be.create_table(tbl["name"], types)
        t.append_pages(pages)
        tables.append(t)
    plan = helpers.plan_from_dict(tables, case["plan"])
    res = be.select(plan)
    got = [helpers.norm_row(r) for r in res.rows()]
    want = [helpers.norm_row(r) for r in case["rows"]]
    # small results come back in the reference's own row order
    assert got == want
    # and the page images decode to the same rows through the reference's row format
    cells, nulls = capi.unpack_pages(res.fetch_pages(), res.ncols)
    types = res.types
    dec = []
    for r in range(cells.shape[0]):
        dec.append(tuple(None if nulls[r, c] else (float(cells[r, c].view(np.float64)) if types[c] == D else int(cells[r, c]))
                         for c in range(res.ncols)))
    assert dec == want
    res.free()
    for t in tables:
        t.drop()


def test_unpack_roundtrip_and_tombstones(be):
    rng = np.random.default_rng(3)
    n = 1000
    types = [I, D, CT_TINYINT, I]
    cells = np.stack([rng.integers(-2**50, 2**50, n), rng.random(n).view(np.int64), rng.integers(0, 2, n),
                      rng.integers(0, 100, n)], axis=1).astype(np.int64)
    nulls = (rng.random((n, 4)) < 0.15).astype(np.uint8)
    deleted = (rng.random(n) < 0.1).astype(np.uint8)
    pages = capi.pack_pages(types, cells, nulls, deleted)
    t = be.create_table("T", types)
    t.append_pages(pages[:3])
    t.append_pages(pages[3:])  # second append continues at a page boundary
    rs = capi.row_size_of(types)
    rpp = 4095 // rs
    assert t.slots == pages.shape[0] * rpp
    assert t.live_rows() == int((deleted == 0).sum())
    for c in range(4):
        got, valid = t.read_column(c, 0, n)
        want_valid = ((deleted == 0) & (nulls[:, c] == 0)).astype(np.uint8)
        assert np.array_equal(valid, want_valid)
        assert np.array_equal(got.view(np.int64)[want_valid == 1], cells[want_valid == 1, c])
    # tombstone through the (page, slot) API, like executor_delete.c:430
    victims = np.flatnonzero(deleted == 0)[:50]
    t.tombstone(victims // rpp, victims % rpp)
    assert t.live_rows() == int((deleted == 0).sum()) - 50
    t.drop()


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("flags", [0, PLAN_NO_FASTPATH])
def test_readme_query_random(be, seed, flags):
    """README query on random keys with duplicates and NULLs, general operators and auto-selected path"""
    rng = np.random.default_rng(200 + seed)
    na, nb, dom = [(50, 80, 20), (5000, 7000, 900), (40000, 30000, 100000), (3000, 3000, 3000)][seed]
    a, b = rng.integers(-dom // 2, dom // 2, na), rng.integers(-dom // 2, dom // 2, nb)
    an, bn = (rng.random(na) < 0.05).astype(np.uint8), (rng.random(nb) < 0.05).astype(np.uint8)
    ga, oa = both_tables(be, [I], [a], [an], paged=(seed % 2 == 0))
    gb, ob = both_tables(be, [I], [b], [bn], paged=(seed % 2 == 0))
    grows, orows, pages, st = run_both(be, [ga, gb], [oa, ob], flags=flags, joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                       out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    assert grows == orows  # same rows in the same (reference) order
    assert st.kernel_launches > 0
    for t in (ga, gb):
        t.drop()


def test_general_operators_random(be):
    """joins (2- and 3-way), WHERE programs, every aggregate, NULLs: CUDA general path == oracle"""
    rng = np.random.default_rng(11)
    n = 20000
    a_cols = [rng.integers(0, 500, n), (rng.random(n) * 100).round(3)]
    a_nulls = [None, (rng.random(n) < 0.1).astype(np.uint8)]
    b_cols = [rng.integers(0, 500, 3000), rng.integers(-1000, 1000, 3000)]
    b_nulls = [(rng.random(3000) < 0.05).astype(np.uint8), (rng.random(3000) < 0.1).astype(np.uint8)]
    c_cols = [rng.permutation(500)[:400].astype(np.int64), rng.integers(0, 50, 400)]
    ga, oa = both_tables(be, [I, D], a_cols, a_nulls)
    gb, ob = both_tables(be, [I, I], b_cols, b_nulls, paged=True)
    gc, oc = both_tables(be, [I, I], c_cols)

    # GROUP BY with every aggregate
    grows, orows, _, _ = run_both(be, [ga], [oa], group=[(0, 0)],
                                  out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_COUNT_COL, 0, 1), (OUT_SUM, 0, 1),
                                       (OUT_MIN, 0, 1), (OUT_MAX, 0, 1), (OUT_AVG, 0, 1)])
    assert helpers.canon_close(grows, orows, rel=1e-9)

    # 3-way join + WHERE + GROUP BY SUM/AVG (config 4 shape)
    kw = dict(joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
              pred=[("col", 0, 1), ("dbl", 25.0), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
              group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_SUM, 0, 1), (OUT_AVG, 2, 1), (OUT_COUNT_STAR,)])
    grows, orows, _, _ = run_both(be, [ga, gb, gc], [oa, ob, oc], **kw)
    assert len(orows) > 10
    assert helpers.canon_close(grows, orows, rel=1e-9)

    # plain 2-way join with projection, OR/XOR/IN/NOT IN/IS NULL programs; row order must equal the oracle's
    preds = [
        [("col", 1, 1), ("int", 0), ("cmp", 1), ("col", 0, 0), ("int", 400), ("cmp", 6), ("or",)],
        [("col", 1, 1), ("int", 0), ("cmp", 2), ("col", 0, 1), ("dbl", 50.0), ("cmp", 2), ("xor",)],
        [("col", 0, 0), ("int", 3), ("int", 5), ("int", 8), ("int", 13), ("in", 4)],
        [("col", 0, 0), ("int", 3), ("int", 5), ("notin", 2), ("col", 1, 1), ("isnull",), ("and",)],
        [("col", 0, 1), ("isnotnull",), ("col", 0, 0), ("col", 1, 1), ("cmp", 3), ("and",)],
    ]
    for pred in preds:
        grows, orows, pages, _ = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], pred=pred,
                                          out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 0, 1), (OUT_COLUMN, 1, 1)])
        assert len(orows) > 0
        assert [helpers.norm_row(r) for r in grows] == [helpers.norm_row(r) for r in orows]

    # aggregates without GROUP BY over a join; and the empty case (no row, like the reference)
    grows, orows, _, _ = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))],
                                  out=[(OUT_COUNT_STAR,), (OUT_SUM, 1, 1), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1), (OUT_AVG, 1, 1)])
    assert helpers.canon_close(grows, orows, rel=1e-9)
    grows, orows, pages, _ = run_both(be, [ga], [oa], pred=[("col", 0, 0), ("int", 10**6), ("cmp", 2)], out=[(OUT_COUNT_STAR,)])
    assert grows == orows == []
    assert pages.shape[0] == 1 and pages[0, 0] == 1  # one page of empty slots, never zero pages
    # cross join (comma list)
    gs, os_ = both_tables(be, [I], [np.arange(7, dtype=np.int64)])
    grows, orows, _, _ = run_both(be, [gs, gc], [os_, oc], joins=["cross"], out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 1, 1)])
    assert [helpers.norm_row(r) for r in grows] == [helpers.norm_row(r) for r in orows]
    for t in (ga, gb, gc, gs):
        t.drop()


@pytest.mark.parametrize("sel", [0.01, 0.5, 1.0])
def test_scan_filter_aggregate_fastpath(be, sel):
    """config 2 shape: SELECT COUNT(*), SUM(v) ... WHERE k >= lo AND k <= hi - fused kernel vs oracle"""
    rng = np.random.default_rng(21)
    n = 300001  # odd on purpose: exercises the scalar tail
    k = rng.integers(0, 2**31, n)
    v = rng.random(n)
    vn = (rng.random(n) < 0.02).astype(np.uint8)
    m = rng.integers(-2**40, 2**40, n)
    g, o = both_tables(be, [I, D, I], [k, v, m], [None, vn, None])
    lo = 2**29
    hi = lo + int(sel * 2**31) - 1
    pred = [("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)]
    out = [(OUT_COUNT_STAR,), (OUT_SUM, 0, 1), (OUT_COUNT_COL, 0, 1), (OUT_AVG, 0, 1), (OUT_MIN, 0, 2), (OUT_MAX, 0, 2), (OUT_SUM, 0, 2)]
    grows, orows, _, st = run_both(be, [g], [o], pred=pred, out=out)
    assert st.path == capi.PATH_SCAN_AGG
    assert len(grows) == 1
    assert helpers.rows_close(grows, [helpers.norm_row(r) for r in orows], rel=1e-9)
    # integer aggregates are bit-exact
    assert grows[0][0] == orows[0][0] and grows[0][4:] == tuple(orows[0][4:])
    # same answer from the general operators
    g2, _, _, st2 = run_both(be, [g], [o], flags=PLAN_NO_FASTPATH, pred=pred, out=out)
    assert st2.path == capi.PATH_GENERAL
    assert helpers.rows_close(g2, grows, rel=1e-9)
    # double range predicate, yoda form
    pred = [("dbl", 0.25), ("col", 0, 1), ("cmp", 5), ("col", 0, 1), ("dbl", 0.75), ("cmp", 1), ("and",)]
    grows, orows, _, st = run_both(be, [g], [o], pred=pred, out=[(OUT_COUNT_STAR,), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1)])
    assert st.path == capi.PATH_SCAN_AGG
    assert grows == [helpers.norm_row(r) for r in orows]
    g.drop()


@pytest.mark.parametrize("shape", ["uniform", "sparse_overlap", "nulls_paged"])
def test_radix_joincount_fastpath(be, shape):
    """README query at 2^20-2^21 rows: radix path == oracle, bit-exact after canonical ordering"""
    rng = np.random.default_rng(31)
    if shape == "uniform":
        na = nb = 1 << 20
        a, b = rng.integers(0, 1 << 20, na), rng.integers(0, 1 << 20, nb)
        an = bn = None
    elif shape == "sparse_overlap":
        na, nb = 700000, 900001
        a = rng.integers(-5_000_000, 40_000_000, na)
        b = rng.integers(30_000_000, 90_000_000, nb)
        an = bn = None
    else:
        na, nb = 600000, 500000
        a, b = rng.integers(1000, 300000, na), rng.integers(0, 250000, nb)
        an, bn = (rng.random(na) < 0.03).astype(np.uint8), (rng.random(nb) < 0.03).astype(np.uint8)
    paged = shape == "nulls_paged"
    ga, oa = both_tables(be, [I], [a], None if an is None else [an], paged=paged)
    gb, ob = both_tables(be, [I], [b], None if bn is None else [bn], paged=paged)
    kw = dict(joins=[((0, 0), (1, 0))], group=[(1, 0)], out=[(OUT_COUNT_STAR,), (OUT_COLUMN, 0, 0), (OUT_COLUMN, 1, 0)])
    grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], **kw)
    assert st.path == capi.PATH_RADIX_JOINCOUNT
    assert helpers.canon(grows) == helpers.canon(orows)
    # independent cross-check of the invariant: sum of counts == join cardinality
    ua, ca = np.unique(a[an == 0] if an is not None else a, return_counts=True)
    ub, cb = np.unique(b[bn == 0] if bn is not None else b, return_counts=True)
    common, ia, ib = np.intersect1d(ua, ub, return_indices=True)
    assert sum(r[0] for r in grows) == int((ca[ia] * cb[ib]).sum())
    assert len(grows) == common.size
    for t in (ga, gb):
        t.drop()


@pytest.mark.parametrize("order", ["ascending", "descending", "clustered", "sorted_duplicates", "sorted_with_nulls"])
def test_radix_joincount_ordered_keys(be, order):
    """auto-increment style ids (all keys of a tile fall into one or two partitions).  A sorted side without NULLs skips
    pass 1 and is counted straight from the column; with NULLs it is not eligible and, after the skew flag of pass 1,
    the direct-count path answers.  Either way the result equals the oracle's"""
    rng = np.random.default_rng(43)
    n = 1 << 20
    an = None
    if order == "ascending":
        a, b = np.arange(n, dtype=np.int64) + 17, np.arange(n, dtype=np.int64) * 2
    elif order == "descending":  # only non-decreasing columns skip pass 1
        a, b = np.arange(n, dtype=np.int64) + 17, (np.arange(n, dtype=np.int64)[::-1] * 2).copy()
    elif order == "clustered":
        a = np.sort(rng.integers(0, 1 << 21, n))
        b = rng.integers(0, 1 << 21, n)
    elif order == "sorted_duplicates":
        a = np.sort(rng.integers(-(1 << 19), 1 << 19, n))  # about two rows per key, negative keys
        b = np.sort(rng.integers(-(1 << 18), 1 << 20, n + 12345))
    else:
        a = np.arange(n, dtype=np.int64)
        b = rng.integers(0, n, n)
        an = (rng.random(n) < 0.01).astype(np.uint8)
    ga, oa = both_tables(be, [I], [a], None if an is None else [an])
    gb, ob = both_tables(be, [I], [b])
    grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                   out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    if order in ("sorted_with_nulls", "descending"):
        assert st.path in (capi.PATH_RADIX_JOINCOUNT, capi.PATH_DIRECT_COUNT)
    else:
        assert st.path == capi.PATH_RADIX_JOINCOUNT
    assert helpers.canon(grows) == helpers.canon(orows)
    # the table changes: the cached "sorted" verdict must not survive
    if order == "ascending":
        ga.append_columns([np.array([5, 3, 1], dtype=np.int64)])
        oa.append_columns([np.array([5, 3, 1], dtype=np.int64)])
        grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                       out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
        assert helpers.canon(grows) == helpers.canon(orows)
    for t in (ga, gb):
        t.drop()


def test_radix_joincount_heavy_key_falls_back(be):
    """more than 255 equal keys wrap a byte counter: detected by the checksum, redone by the direct-count path (one 32-bit
    counter per key value; no joined pair is materialised there either)"""
    rng = np.random.default_rng(41)
    n = 1 << 20
    a = rng.integers(0, 1 << 20, n)
    b = rng.integers(0, 1 << 20, n)
    a[:5000] = 777  # heavy hitter
    b[:300] = 777
    ga, oa = both_tables(be, [I], [a])
    gb, ob = both_tables(be, [I], [b])
    grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                   out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    assert st.path == capi.PATH_DIRECT_COUNT
    assert helpers.canon(grows) == helpers.canon(orows)
    for t in (ga, gb):
        t.drop()


def test_generate_is_sharding_invariant_and_join_properties(be):
    """device-side generator: any sharding of (seed, global index) gives the same table; at 2^24 x 2^24 the
    radix path satisfies the domain invariants (groups <= min side, keys sorted-unique, sum(count) == |A join B|)"""
    spec = capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 24) - 1, seed=3)
    whole = be.create_table("w", [I])
    whole.generate(1 << 16, [spec])
    parts = be.create_table("p", [I])
    parts.generate(1 << 15, [spec], row_offset=0)
    parts.generate(1 << 15, [spec], row_offset=1 << 15)
    w, _ = whole.read_column(0)
    p, _ = parts.read_column(0)
    assert np.array_equal(w, p)
    assert w.min() >= 0 and w.max() < (1 << 24)
    whole.drop()
    parts.drop()

    n = 1 << 24
    ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
    ta.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=1)])
    tb.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=2)])
    plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    res = be.select(plan)
    assert be.stats().path == capi.PATH_RADIX_JOINCOUNT
    (keys, cnts), _ = res.fetch_columns()
    res.free()
    a, _ = ta.read_column(0)
    b, _ = tb.read_column(0)
    ca, cb = np.bincount(a, minlength=n), np.bincount(b, minlength=n)
    want_keys = np.flatnonzero((ca > 0) & (cb > 0))
    order = np.argsort(keys)
    assert np.array_equal(keys[order], want_keys)
    assert np.array_equal(cnts[order], (ca * cb)[want_keys])
    ta.drop()
    tb.drop()


@pytest.mark.parametrize("fact_first", [False, True])
def test_star_join_fastpath(be, fact_first):
    """BASELINE config 5 shape: small dimension with unique dense keys, large fact table, GROUP BY a dimension column
    with MIN / MAX / SUM / COUNT over fact columns: direct-addressed star join == oracle (ints bit-exact, DOUBLE SUM 1e-9)"""
    rng = np.random.default_rng(51)
    nd, nf = 60000, (1 << 20) + 4321
    d_id = (rng.permutation(65536)[:nd] + 1000).astype(np.int64)        # unique, some keys of the range missing
    d_g = rng.integers(-7, 1017, nd)
    f_fk = rng.integers(900, 1000 + 65536 + 100, nf)                     # some foreign keys match nothing
    f_m = rng.integers(-(1 << 40), 1 << 40, nf)
    f_x = rng.random(nf) * 100.0
    f_fk_null = (rng.random(nf) < 0.01).astype(np.uint8)
    f_m_null = (rng.random(nf) < 0.05).astype(np.uint8)
    gd, od = both_tables(be, [I, I], [d_id, d_g])
    gf, of = both_tables(be, [I, I, D], [f_fk, f_m, f_x], [f_fk_null, f_m_null, None])
    if fact_first:
        gt, ot, d, f = [gf, gd], [of, od], 1, 0
    else:
        gt, ot, d, f = [gd, gf], [od, of], 0, 1
    # (the ON operands name the earlier table first: column 0 is the key of both tables)
    kw = dict(joins=[((0, 0), (1, 0))], group=[(d, 1)],
              out=[(OUT_COLUMN, d, 1), (OUT_MIN, f, 1), (OUT_MAX, f, 1), (OUT_COUNT_STAR,), (OUT_SUM, f, 2)])
    grows, orows, _, st = run_both(be, gt, ot, **kw)
    assert st.path == capi.PATH_DIRECT_STAR
    assert len(orows) > 1000
    assert helpers.canon_close(grows, orows, rel=1e-9)
    # AVG and COUNT(col) through the same kernel
    kw["out"] = [(OUT_AVG, f, 1), (OUT_COLUMN, d, 1), (OUT_COUNT_COL, f, 1), (OUT_AVG, f, 2)]
    grows, orows, _, st = run_both(be, gt, ot, **kw)
    assert st.path == capi.PATH_DIRECT_STAR
    assert helpers.canon_close(grows, orows, rel=1e-9)
    # a dimension with a duplicate key is not a star dimension: the general operators answer, same result as the oracle
    gd.append_columns([d_id[:3].copy(), np.array([1, 2, 3], dtype=np.int64)])
    od.append_columns([d_id[:3].copy(), np.array([1, 2, 3], dtype=np.int64)])
    kw["out"] = [(OUT_COLUMN, d, 1), (OUT_MIN, f, 1), (OUT_COUNT_STAR,)]
    grows, orows, _, st = run_both(be, gt, ot, **kw)
    assert st.path == capi.PATH_GENERAL
    assert helpers.canon_close(grows, orows, rel=1e-9)
    for t in (gd, gf):
        t.drop()


@pytest.mark.parametrize("mode", ["dense", "dense_negative_nulls", "hash_wide", "hash_double", "hash_composite"])
def test_group_by_slot_modes(be, mode):
    """the hash aggregate's slot schemes (direct slots for a narrow INT key, hash table otherwise), each with a WHERE whose
    verdict bitmap is consumed by the aggregate, > 2048 groups so that the per-CTA group cache has conflicts, NULL group,
    INT64_MIN key"""
    rng = np.random.default_rng(77)
    n = 30000
    v = (rng.random(n) * 1000).round(2)
    vn = (rng.random(n) < 0.1).astype(np.uint8)
    w = rng.integers(-50, 50, n)
    kn = None
    types = [I, D, I]
    group = [(0, 0)]
    if mode == "dense":
        k = rng.integers(0, 5000, n)
    elif mode == "dense_negative_nulls":
        k = rng.integers(-3000, 3000, n)
        kn = (rng.random(n) < 0.05).astype(np.uint8)
    elif mode == "hash_wide":
        k = rng.integers(-2500, 2500, n) * (10**12)
        k[:7] = np.iinfo(np.int64).min
        kn = (rng.random(n) < 0.05).astype(np.uint8)
    elif mode == "hash_double":
        types = [D, D, I]
        k = rng.integers(0, 4000, n) / 8.0
        k[:5] = -0.0
        k[5:9] = 0.0
    else:
        k = rng.integers(0, 70, n)
        group = [(0, 0), (0, 2)]
    gt, ot = both_tables(be, types, [k, v, w], [kn, vn, None])
    out = [(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_COUNT_COL, 0, 1), (OUT_SUM, 0, 1), (OUT_MIN, 0, 2), (OUT_MAX, 0, 1), (OUT_AVG, 0, 2)]
    if mode == "hash_composite":
        out.insert(1, (OUT_COLUMN, 0, 2))
    for pred in (None, [("col", 0, 2), ("int", 10), ("cmp", 1), ("col", 0, 1), ("dbl", 900.0), ("cmp", 1), ("and",)]):
        kw = dict(group=group, out=out)
        if pred:
            kw["pred"] = pred
        grows, orows, _, st = run_both(be, [gt], [ot], flags=PLAN_NO_FASTPATH, **kw)
        assert len(orows) > 2048
        assert helpers.canon_close(grows, orows, rel=1e-9)
    gt.drop()


@pytest.mark.parametrize("shape", ["all_match", "some_miss", "none_match", "null_keys", "wide_keys", "wide_dups", "double_keys"])
def test_join_with_unique_build_side(be, shape):
    """foreign-key joins (build side without duplicate keys) append a row-id column instead of expanding the tuples;
    the tuples are compacted only when some tuple has no partner. Row order must stay the reference's.
    Narrow INT keys take the direct row table, wide INT keys and DOUBLE keys the hash table, duplicates the CSR route."""
    rng = np.random.default_rng(91)
    n, nd = 6000, 300
    fk = rng.integers(0, nd, n)
    fkn = (rng.random(n) < 0.08).astype(np.uint8) if shape == "null_keys" else None
    val = rng.integers(-100, 100, n)
    dk = rng.permutation(nd).astype(np.int64)
    if shape == "some_miss":
        dk = dk[: nd // 2]
    elif shape == "none_match":
        dk = dk + 10 * nd
    kt = I
    if shape in ("wide_keys", "wide_dups"):
        fk, dk = fk * 10**10 - 7, dk * 10**10 - 7
        if shape == "wide_dups":
            dk = np.concatenate([dk, dk[:40]])
    elif shape == "double_keys":
        kt = D
        fk, dk = fk / 4.0, dk / 4.0
    dv = rng.integers(0, 9, len(dk))
    dkn = (rng.random(len(dk)) < 0.1).astype(np.uint8) if shape == "null_keys" else None
    ek = rng.permutation(nd).astype(np.int64)  # second dimension: always complete
    ev = (rng.random(nd) * 10).round(1)
    gf, of = both_tables(be, [kt, I, I], [fk, val, rng.integers(0, nd, n)], [fkn, None, None])
    gd, od = both_tables(be, [kt, I], [dk, dv], [dkn, None], paged=True)
    ge, oe = both_tables(be, [I, D], [ek, ev])
    # projection of a 3-way join, in the reference's row order
    grows, orows, _, _ = run_both(be, [gf, gd, ge], [of, od, oe], flags=PLAN_NO_FASTPATH, joins=[((0, 0), (1, 0)), ((0, 2), (2, 0))],
                                  out=[(OUT_COLUMN, 0, 1), (OUT_COLUMN, 1, 1), (OUT_COLUMN, 2, 1)])
    assert (len(orows) == 0) == (shape == "none_match")
    assert [helpers.norm_row(r) for r in grows] == [helpers.norm_row(r) for r in orows]
    # and under WHERE + GROUP BY
    grows, orows, _, _ = run_both(be, [gf, gd, ge], [of, od, oe], flags=PLAN_NO_FASTPATH, joins=[((0, 0), (1, 0)), ((0, 2), (2, 0))],
                                  pred=[("col", 0, 1), ("int", 50), ("cmp", 1)], group=[(1, 1)],
                                  out=[(OUT_COLUMN, 1, 1), (OUT_COUNT_STAR,), (OUT_SUM, 0, 1), (OUT_AVG, 2, 1)])
    assert helpers.canon_close(grows, orows, rel=1e-9)
    for t in (gf, gd, ge):
        t.drop()


@pytest.mark.parametrize("shape", ["all_match", "some_miss", "none_match", "null_keys", "snowflake", "tombstones", "wide_keys", "dups"])
def test_fused_multiway_aggregate(be, shape, monkeypatch):
    """default path for star / snowflake aggregates (MDBCU_FUSED_MULTIWAY=0 disables it): joins against duplicate-free narrow INT keys + WHERE + aggregates in one kernel
    over tables[0], no tuple arrays. Same answers as the oracle; plans outside the shape fall through to the general operators."""
    monkeypatch.setenv("MDBCU_FUSED_MULTIWAY", "1")
    rng = np.random.default_rng(131)
    n, nd = 20000, 700
    fk = rng.integers(0, nd, n)
    fk2 = rng.integers(0, nd, n)
    val = (rng.random(n) * 100).round(2)
    valn = (rng.random(n) < 0.1).astype(np.uint8)
    fkn = (rng.random(n) < 0.08).astype(np.uint8) if shape == "null_keys" else None
    dk = rng.permutation(nd).astype(np.int64)
    if shape == "some_miss":
        dk = dk[: nd // 2]
    elif shape == "none_match":
        dk = dk + 10 * nd
    elif shape == "wide_keys":
        fk, dk = fk * 10**10, dk * 10**10
    elif shape == "dups":
        dk = np.concatenate([dk, dk[:25]])
    dv = rng.integers(0, 40, len(dk))
    dlink = rng.integers(0, nd, len(dk))  # snowflake: the second dimension hangs off the first
    dkn = (rng.random(len(dk)) < 0.1).astype(np.uint8) if shape == "null_keys" else None
    ek = rng.permutation(nd).astype(np.int64)
    ev = (rng.random(nd) * 10).round(1)
    if shape == "tombstones":
        types = [I, I, D]
        cells = np.stack([fk, fk2, val.view(np.int64)], axis=1)
        nulls = np.stack([np.zeros(n, np.uint8), np.zeros(n, np.uint8), valn], axis=1)
        pages = capi.pack_pages(types, cells, nulls, (rng.random(n) < 0.2).astype(np.uint8))
        gf, of = be.create_table("f", types), oracle.OracleTable(types)
        gf.append_pages(pages)
        of.append_pages(pages)
    else:
        gf, of = both_tables(be, [I, I, D], [fk, fk2, val], [fkn, None, valn])
    gd, od = both_tables(be, [I, I, I], [dk, dv, dlink], [dkn, None, None], paged=True)
    ge, oe = both_tables(be, [I, D], [ek, ev])
    second = ((1, 2), (2, 0)) if shape == "snowflake" else ((0, 1), (2, 0))
    joins = [((0, 0), (1, 0)), second]
    fused = shape not in ("wide_keys", "dups")
    pred = [("col", 0, 2), ("dbl", 20.0), ("cmp", 6), ("col", 1, 1), ("int", 30), ("cmp", 1), ("and",)]
    plans = [
        dict(group=[(1, 1)], out=[(OUT_COLUMN, 1, 1), (OUT_COUNT_STAR,), (OUT_SUM, 0, 2), (OUT_AVG, 2, 1), (OUT_MIN, 0, 2), (OUT_MAX, 2, 1)]),
        dict(group=[(1, 1)], pred=pred, out=[(OUT_COLUMN, 1, 1), (OUT_COUNT_COL, 0, 2), (OUT_SUM, 0, 2), (OUT_AVG, 2, 1)]),
        dict(pred=pred, out=[(OUT_COUNT_STAR,), (OUT_SUM, 0, 2), (OUT_MAX, 2, 1)]),
        dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 1, 1), (OUT_COLUMN, 2, 1), (OUT_COUNT_STAR,)]),
        dict(group=[(0, 0), (0, 1)], pred=pred, out=[(OUT_COLUMN, 0, 0), (OUT_COLUMN, 0, 1), (OUT_SUM, 0, 2)]),
    ]
    if shape in ("null_keys", "wide_keys"):
        plans.pop()  # composite GROUP BY over a column with NULLs or outside the int32 range is rejected by every path
    for kw in plans:
        grows, orows, _, st = run_both(be, [gf, gd, ge], [of, od, oe], flags=PLAN_NO_FASTPATH, joins=joins, **kw)
        assert st.path == (capi.PATH_FUSED_MULTIWAY if fused else capi.PATH_GENERAL)
        assert (len(orows) == 0) == (shape == "none_match")
        assert helpers.canon_close(grows, orows, rel=1e-9)
    # a projection (no aggregate) never takes the fused path
    grows, orows, _, st = run_both(be, [gf, gd, ge], [of, od, oe], flags=PLAN_NO_FASTPATH, joins=joins,
                                   out=[(OUT_COLUMN, 0, 2), (OUT_COLUMN, 1, 1)])
    assert st.path == capi.PATH_GENERAL
    assert [helpers.norm_row(r) for r in grows] == [helpers.norm_row(r) for r in orows]
    for t in (gf, gd, ge):
        t.drop()


def zipf_keys(rng, n, domain, s=1.1):
    """Zipf(s) over [0, domain): rank r has weight 1 / (r + 1)^s (BASELINE configs[3])"""
    w = 1.0 / np.arange(1, domain + 1) ** s
    return rng.choice(domain, size=n, p=w / w.sum()).astype(np.int64)


@pytest.mark.parametrize("fused", [False, True])
def test_zipf_three_way_join_group_cache(be, fused, monkeypatch):
    """config 4 shape at test scale: Zipf(1.1) foreign keys (12 % of the rows in one group: the per-CTA cache of hot groups
    and the global accumulators both get traffic), two unique dimensions, WHERE on fact and dimension, SUM / AVG / COUNT.
    General operators and the fused multiway kernel against the oracle: keys and counts exact, DOUBLE SUM / AVG 1e-9."""
    monkeypatch.setenv("MDBCU_FUSED_MULTIWAY", "1" if fused else "0")
    rng = np.random.default_rng(404)
    n, nd = 1 << 18, 1 << 13
    a_id = zipf_keys(rng, n, nd)
    assert np.bincount(a_id).max() > n // 20  # the hot group really is hot
    x = rng.random(n)
    ga, oa = both_tables(be, [I, D], [a_id, x])
    gb, ob = both_tables(be, [I, I], [rng.permutation(nd).astype(np.int64), rng.integers(0, 1000, nd)])
    gc, oc = both_tables(be, [I, I], [rng.permutation(nd).astype(np.int64), rng.integers(0, 50, nd)])
    kw = dict(joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
              pred=[("col", 0, 1), ("dbl", 0.25), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
              group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_SUM, 0, 1), (OUT_AVG, 2, 1), (OUT_COUNT_STAR,)])
    grows, orows, _, st = run_both(be, [ga, gb, gc], [oa, ob, oc], **kw)
    assert st.path == (capi.PATH_FUSED_MULTIWAY if fused else capi.PATH_GENERAL)
    assert len(orows) > 1000
    assert helpers.canon_close(grows, orows, rel=1e-9)
    # keys and counts are bit-exact
    assert sorted((r[0], r[3]) for r in grows) == sorted((r[0], r[3]) for r in orows)
    # Zipf keys on BOTH sides of the README query: heavy duplicates, the radix path must hand over or answer exactly
    # (numpy histograms are the check: the oracle's hash join would visit every one of the 3 * 10^10 joined pairs)
    za, zb = zipf_keys(rng, 1 << 20, 1 << 16), zipf_keys(rng, 1 << 20, 1 << 16)
    gz, gy = be.create_table("z", [I]), be.create_table("y", [I])
    gz.append_columns([za])
    gy.append_columns([zb])
    res = be.select(capi.make_plan([gz, gy], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)]))
    assert be.stats().path == capi.PATH_DIRECT_COUNT  # no joined pair is materialised
    (keys, cnts), _ = res.fetch_columns()
    res.free()
    prod = np.bincount(za, minlength=1 << 16) * np.bincount(zb, minlength=1 << 16)
    assert np.array_equal(prod[keys], cnts) and keys.size == np.count_nonzero(prod) == np.unique(keys).size
    assert int(cnts.sum()) > 1 << 32
    for t in (ga, gb, gc, gz, gy):
        t.drop()


def test_scan_aggregate_double_sum_2p24(be):
    """the 1e-9 claim for DOUBLE SUM / AVG at 2^24 rows (the block partials are folded in a fixed order): against the
    oracle's sequential sum and against numpy's pairwise sum; COUNT / MIN / MAX bit-exact"""
    rng = np.random.default_rng(2024)
    n = 1 << 24
    k = rng.integers(0, 2**31, n)
    v = rng.random(n) * 1e6 - 3e5  # mixed signs: cancellation makes the tolerance meaningful
    g, o = both_tables(be, [I, D], [k, v])
    lo, hi = 2**29, 3 * 2**29 - 1
    pred = [("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)]
    out = [(OUT_COUNT_STAR,), (OUT_SUM, 0, 1), (OUT_AVG, 0, 1), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1)]
    grows, orows, _, st = run_both(be, [g], [o], pred=pred, out=out)
    assert st.path == capi.PATH_SCAN_AGG
    assert helpers.rows_close(grows, [helpers.norm_row(r) for r in orows], rel=1e-9)
    m = (k >= lo) & (k <= hi)
    want_sum = float(np.sum(v[m]))
    assert grows[0][0] == int(m.sum())
    assert abs(grows[0][1] - want_sum) <= 1e-9 * abs(want_sum)
    assert grows[0][3] == float(v[m].min()) and grows[0][4] == float(v[m].max())
    # two runs give the same bits (deterministic reduction order)
    g2, _, _, _ = run_both(be, [g], [o], pred=pred, out=out)
    assert g2 == grows
    g.drop()


@pytest.mark.parametrize("log2_rows", [24, 26])
def test_radix_joincount_large_against_numpy(be, log2_rows):
    """README query at 2^24 and 2^26 rows per side, generated on the device: every (key, count) row against numpy histograms of
    the mirrored columns (independent of the library and of the oracle), all keys accounted for"""
    n = 1 << log2_rows
    ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
    ta.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=31)])
    tb.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=32)])
    res = be.select(capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)]))
    assert be.stats().path == capi.PATH_RADIX_JOINCOUNT
    (keys, cnts), _ = res.fetch_columns()
    res.free()
    a, _ = ta.read_column(0)
    b, _ = tb.read_column(0)
    expect = np.bincount(a, minlength=n) * np.bincount(b, minlength=n)
    assert np.array_equal(expect[keys], cnts) and cnts.min() >= 1
    assert np.unique(keys).size == keys.size == np.count_nonzero(expect)
    assert int(cnts.sum()) == int(expect.sum())  # = the join's cardinality
    ta.drop()
    tb.drop()


@pytest.mark.parametrize("where", ["across_2p32", "across_zero", "above_2p33", "negative"])
@pytest.mark.parametrize("layout", ["key_count", "count_key"])
def test_radix_joincount_key_words_and_layouts(be, where, layout):
    """pass 2 writes keys as (high word of the partition, low word + offset) and addresses rows with 32-bit arithmetic when
    the partition allows it: partitions that straddle a multiple of 2^32 (or zero), keys far above 2^32 and negative keys
    must come out exactly, in both compiled result layouts"""
    rng = np.random.default_rng(97)
    n = 1 << 20
    base = {"across_2p32": (1 << 32) - (1 << 19), "across_zero": -(1 << 19), "above_2p33": (1 << 33) + 12345,
            "negative": -(1 << 40) - 999}[where]
    a = base + rng.integers(0, 1 << 20, n)
    b = base + rng.integers(0, 1 << 20, n + 777)
    ga, oa = both_tables(be, [I], [a])
    gb, ob = both_tables(be, [I], [b])
    out = [(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)] if layout == "key_count" else [(OUT_COUNT_STAR,), (OUT_COLUMN, 1, 0)]
    grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=out)
    assert st.path == capi.PATH_RADIX_JOINCOUNT
    assert helpers.canon(grows) == helpers.canon(orows)
    for t in (ga, gb):
        t.drop()


@pytest.mark.parametrize("log2_half_range", [28, 31])
def test_join_count_wide_key_range_takes_direct_count(be, log2_half_range):
    """keys spread over more than 2^28 values (4096 partitions x 2^16 remainders): the radix path hands the query to the
    direct-count path (one 32-bit counter per key value), not to the general operators; a range of 2^32 values - ordinary
    32-bit keys - is counted in four windows of 2^30"""
    rng = np.random.default_rng(101)
    n = 1 << 21
    h = 1 << log2_half_range
    a = rng.integers(-h, h, n)
    b = rng.integers(-h, h, n + 4321)
    b[:50000] = a[:50000]  # make sure there are matches
    a[:3] = [-h, h - 1, 0]
    b[-3:] = [-h, h - 1, 0]
    ga, oa = both_tables(be, [I], [a])
    gb, ob = both_tables(be, [I], [b])
    grows, orows, _, st = run_both(be, [ga, gb], [oa, ob], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                   out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    assert st.path == capi.PATH_DIRECT_COUNT
    assert len(grows) >= 50000
    assert helpers.canon(grows) == helpers.canon(orows)
    for t in (ga, gb):
        t.drop()


@pytest.mark.parametrize("n", [1000, 4097, 300001])
def test_predicate_scan_term_programs(be, n):
    """K1 on one fully live table (identity tuples, k_eval_terms_scan): WHERE programs that are AND / OR / XOR combinations of
    column-vs-literal and IS [NOT] NULL terms are compiled into a truth table on the host; programs outside that shape (IN
    lists, column vs column, more than eight terms) take the postfix interpreter.  Projection and aggregate consumers, NULL
    cells, DOUBLE comparisons, literals on the left, ragged sizes (not a multiple of 256 rows): all equal to the oracle, and
    projections come back in storage order."""
    rng = np.random.default_rng(500 + n)
    cols = [rng.integers(-50, 50, n), rng.random(n), rng.integers(0, 1000, n)]
    nulls = [(rng.random(n) < 0.1).astype(np.uint8), (rng.random(n) < 0.1).astype(np.uint8), np.zeros(n, np.uint8)]
    g, o = both_tables(be, [I, D, I], cols, nulls)
    preds = [
        # (a < 0 OR a >= 40)
        [("col", 0, 0), ("int", 0), ("cmp", 1), ("col", 0, 0), ("int", 40), ("cmp", 6), ("or",)],
        # (0.5 > b) XOR (c <> 7)   - literal on the left, DOUBLE compare
        [("dbl", 0.5), ("col", 0, 1), ("cmp", 2), ("col", 0, 2), ("int", 7), ("cmp", 3), ("xor",)],
        # a IS NULL OR (b IS NOT NULL AND b <= 0.25 AND c >= 100)
        [("col", 0, 0), ("isnull",), ("col", 0, 1), ("isnotnull",), ("col", 0, 1), ("dbl", 0.25), ("cmp", 5), ("and",),
         ("col", 0, 2), ("int", 100), ("cmp", 6), ("and",), ("or",)],
        # a conjunction of three terms (the early-exit shape), an INT column against a DOUBLE literal
        [("col", 0, 0), ("dbl", -10.5), ("cmp", 2), ("col", 0, 2), ("int", 900), ("cmp", 1), ("and",), ("col", 0, 1), ("dbl", 0.9), ("cmp", 1), ("and",)],
        # eight terms: ((t1 OR t2) AND (t3 OR t4)) XOR ((t5 AND t6) OR (t7 AND t8))
        [("col", 0, 0), ("int", -25), ("cmp", 1), ("col", 0, 0), ("int", 25), ("cmp", 2), ("or",),
         ("col", 0, 2), ("int", 500), ("cmp", 1), ("col", 0, 1), ("dbl", 0.75), ("cmp", 6), ("or",), ("and",),
         ("col", 0, 2), ("int", 250), ("cmp", 6), ("col", 0, 2), ("int", 750), ("cmp", 5), ("and",),
         ("col", 0, 1), ("dbl", 0.1), ("cmp", 1), ("col", 0, 0), ("int", 0), ("cmp", 4), ("and",), ("or",), ("xor",)],
        # outside the compiled shape: IN list, column vs column, nine terms
        [("col", 0, 2), ("int", 3), ("int", 5), ("int", 8), ("in", 3), ("col", 0, 0), ("int", 10), ("cmp", 2), ("or",)],
        [("col", 0, 0), ("col", 0, 2), ("cmp", 1), ("col", 0, 1), ("dbl", 0.5), ("cmp", 2), ("and",)],
        [("col", 0, 2), ("int", 100), ("cmp", 1)] + sum([[("col", 0, 2), ("int", 100 + 90 * k), ("cmp", 2), ("xor",)] for k in range(1, 9)], []),
    ]
    for pred in preds:
        grows, orows, _, st = run_both(be, [g], [o], pred=pred, out=[(OUT_COLUMN, 0, 2), (OUT_COLUMN, 0, 0), (OUT_COLUMN, 0, 1)])
        assert st.path == capi.PATH_GENERAL
        assert grows == [helpers.norm_row(r) for r in orows], pred  # same rows, same (storage) order
        grows, orows, _, _ = run_both(be, [g], [o], pred=pred, group=[(0, 0)],
                                      out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_SUM, 0, 2), (OUT_MIN, 0, 1)])
        assert helpers.canon_close(grows, orows, rel=1e-9), pred
    g.drop()


def _tail_cases():
    from tests.test_oracle import TAIL_CASES
    return TAIL_CASES


@pytest.mark.parametrize("case", _tail_cases(), ids=[c[0] for c in _tail_cases()])
def test_tail_operators_match_oracle(be, case):
    """HAVING / DISTINCT / ORDER BY / LIMIT on the device-resident result (mdb_tail.cu) against the oracle, whose semantics
    are pinned to sqlite3 (tests/test_oracle.py); ordered cases are compared row by row, in order"""
    from tests.test_oracle import tail_tables
    name, kw, _, ordered = case
    a_rows, b_rows = tail_tables()

    def cols(rows, dbl):
        k = np.array([r[0] for r in rows], dtype=np.int64)
        v = np.array([0 if r[1] is None else r[1] for r in rows], dtype=np.float64 if dbl else np.int64)
        return [k, v], [None, np.array([r[1] is None for r in rows], dtype=np.uint8)]

    (ca, na), (cb, nb) = cols(a_rows, True), cols(b_rows, False)
    ga, oa = both_tables(be, [I, D], ca, na)
    gb, ob = both_tables(be, [I, I], cb, nb)
    gt, ot = ([ga, gb], [oa, ob]) if kw.get("joins") else ([ga], [oa])
    for flags in (0, PLAN_NO_FASTPATH):
        grows, orows, pages, st = run_both(be, gt, ot, flags=flags, **kw)
        if ordered:
            assert helpers.rows_close([helpers.norm_row(r) for r in grows], [helpers.norm_row(r) for r in orows])
            # the page images carry the same order (what query_cur_step walks)
            cells, nulls = capi.unpack_pages(pages, len(kw["out"]))
            assert cells.shape[0] == len(orows)
        else:
            assert helpers.canon_close(grows, orows)
    for t in (ga, gb):
        t.drop()


def test_tail_operators_large_result(be):
    """ORDER BY / LIMIT / DISTINCT / HAVING on a 2^20-row result of the radix path: top-k by count, then by key"""
    rng = np.random.default_rng(123)
    n = 1 << 20
    a, b = rng.integers(0, n // 4, n), rng.integers(0, n // 4, n)
    ga, gb = be.create_table("A", [I]), be.create_table("B", [I])
    ga.append_columns([a])
    gb.append_columns([b])
    base = dict(joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
    ca, cb = np.bincount(a, minlength=n // 4), np.bincount(b, minlength=n // 4)
    prod = ca * cb
    keys = np.flatnonzero(prod)
    res = be.select(capi.make_plan([ga, gb], order=[(1, True), (0, False)], limit=1000, **base))
    assert be.stats().path == capi.PATH_RADIX_JOINCOUNT
    (k, c), _ = res.fetch_columns()
    res.free()
    order = np.lexsort((keys, -prod[keys]))[:1000]
    assert np.array_equal(k, keys[order]) and np.array_equal(c, prod[keys][order])
    # HAVING COUNT(*) >= 40 ORDER BY key DESC
    res = be.select(capi.make_plan([ga, gb], having=[("out", 1), ("int", 40), ("cmp", 6)], order=[(0, True)], **base))
    (k, c), _ = res.fetch_columns()
    res.free()
    want = keys[prod[keys] >= 40][::-1]
    assert np.array_equal(k, want) and np.array_equal(c, prod[want])
    # DISTINCT over the counts alone
    res = be.select(capi.make_plan([ga, gb], joins=base["joins"], group=base["group"], out=[(OUT_COUNT_STAR,)], distinct=True, order=[(0, False)]))
    (c,), _ = res.fetch_columns()
    res.free()
    assert np.array_equal(c, np.unique(prod[keys]))
    for t in (ga, gb):
        t.drop()
