"""GPU tests (-m gpu) of the drop-in surface: the reference's own executor tests
(tests/engine/executor_select.c:47-401) and more, run VERBATIM as SQL through this repo's
database_open / query_execute / query_cur_step / query_column_int64 / query_free (libmidoridb_b200.so) and read
through the cursor API exactly like the reference's tests do."""
import numpy as np
import pytest

from midoridb_b200 import capi, db as mdb
from tests import helpers

pytestmark = pytest.mark.gpu
CASES = helpers.load_golden()


def _load_case(db, case):
    for tbl in case["tables"]:
        cols = ", ".join("%s %s" % (c, "DOUBLE" if t == capi.CT_DOUBLE else "INT") for c, t in zip(tbl["cols"], tbl["types"]))
        assert db.execute("CREATE TABLE %s (%s);" % (tbl["name"], cols)) == 0
        rows = tbl["rows"]
        for i in range(0, len(rows), 200):
            vals = ", ".join("(" + ", ".join("NULL" if v is None else repr(v) for v in r["v"]) + ")" for r in rows[i:i + 200])
            assert db.execute("INSERT INTO %s VALUES %s;" % (tbl["name"], vals)) == len(rows[i:i + 200])


@pytest.mark.parametrize("case", [c for c in CASES if not c["stmts"]], ids=[c["name"] for c in CASES if not c["stmts"]])
def test_reference_sql_verbatim(case):
    with mdb.Database() as db:
        _load_case(db, case)
        names, rows = db.query(case["sql"])
        assert names == case["columns"]                      # fully-qualified names in the reference's scaffold order
        assert [helpers.norm_row(r) for r in rows] == [helpers.norm_row(r) for r in case["rows"]]  # reference row order
        path, launches = db.last_path()
        assert launches > 0


@pytest.mark.parametrize("case", [c for c in CASES if c["stmts"]], ids=[c["name"] for c in CASES if c["stmts"]])
def test_reference_dml_then_select(case):
    """DELETE before SELECT: the tombstones must reach the device mirror (executor_delete.c:430 hook)"""
    with mdb.Database() as db:
        for tbl in case["tables"]:
            cols = ", ".join("%s INT" % c for c in tbl["cols"])
            db.execute("CREATE TABLE %s (%s);" % (tbl["name"], cols))
            vals = ", ".join("(" + ", ".join("NULL" if v is None else repr(v) for v in r["v"]) + ")" for r in tbl["rows"])
            db.execute("INSERT INTO %s VALUES %s;" % (tbl["name"], vals))
        # the golden tables hold the post-DML state (deleted flags); replay the statements on a warm mirror
        db.query("SELECT %s FROM %s;" % (case["tables"][0]["cols"][0], case["tables"][0]["name"]))
        for s in case["stmts"]:
            db.execute(s)
        names, rows = db.query(case["sql"])
        assert names == case["columns"]
        assert [helpers.norm_row(r) for r in rows] == [helpers.norm_row(r) for r in case["rows"]]


def test_readme_program():
    """the README example (README.md:48-77) end to end"""
    with mdb.Database() as db:
        db.execute("CREATE TABLE A (id_a INT);")
        db.execute("CREATE TABLE B (id_b INT);")
        db.execute("INSERT INTO A VALUES (1),(3),(4);")
        db.execute("INSERT INTO B VALUES (1),(1),(3),(3),(4),(NULL);")
        names, rows = db.query("SELECT     id_a, COUNT(*) FROM     A INNER JOIN B     ON A.id_a = B.id_b GROUP BY     id_a;")
        assert names == ["A.id_a", "COUNT(*)"] and rows == [(1, 2), (3, 2), (4, 1)]


def test_update_insert_after_mirror_and_multipage_cursor():
    rng = np.random.default_rng(5)
    n = 1000  # 8 pages of 127 rows: the reference's cursor emits a bogus row per page boundary (D1); ours must not
    k = rng.integers(0, 1000, n)
    with mdb.Database() as db:
        db.execute("CREATE TABLE T (k INT, v INT);")
        for i in range(0, n, 250):
            db.execute("INSERT INTO T VALUES %s;" % ", ".join("(%d, %d)" % (int(k[j]), j) for j in range(i, i + 250)))
        names, rows = db.query("SELECT k, v FROM T WHERE k >= 250 AND k <= 749;")
        want = [(int(k[j]), j) for j in range(n) if 250 <= k[j] <= 749]
        assert rows == want
        # UPDATE in place, then more INSERTs into the partially filled last page
        assert db.execute("UPDATE T SET v = -1 WHERE k < 10;") == int((k < 10).sum())
        assert db.execute("UPDATE T SET k = NULL WHERE v = 999;") == 1
        db.execute("INSERT INTO T VALUES (5, 5000), (2000, 5001);")
        names, rows = db.query("SELECT k, v FROM T WHERE k < 10 OR k > 1500;")
        exp = [(int(k[j]), -1) for j in range(n) if k[j] < 10 and j != 999] + [(5, 5000), (2000, 5001)]
        assert rows == exp
        names, rows = db.query("SELECT COUNT(*) FROM T WHERE k IS NULL;")
        assert rows == [(1,)]
        assert db.execute("DELETE FROM T WHERE v = -1;") == int(((k < 10) & (np.arange(n) != 999)).sum())
        names, rows = db.query("SELECT COUNT(*) FROM T WHERE k < 10;")
        assert rows == [(1,)]


def test_extensions_and_errors():
    with mdb.Database() as db:
        db.execute("CREATE TABLE T (k INT, v DOUBLE);")
        db.execute("INSERT INTO T VALUES (1, 0.5), (1, 1.5), (2, NULL), (2, 4.0), (3, NULL);")
        names, rows = db.query("SELECT k, SUM(v), AVG(v), MIN(v), MAX(v), COUNT(*) FROM T GROUP BY k;")
        got = {r[names.index("T.k")]: r for r in rows}
        i = {n: names.index(n) for n in names}
        assert got[1][i["SUM(T.v)"]] == 2.0 and got[1][i["AVG(T.v)"]] == 1.0 and got[1][i["COUNT(*)"]] == 2
        assert got[2][i["MIN(T.v)"]] == 4.0 and got[2][i["MAX(T.v)"]] == 4.0
        assert got[3][i["SUM(T.v)"]] is None and got[3][i["COUNT(*)"]] == 1
        names, rows = db.query("SELECT COUNT(*) FROM T WHERE k BETWEEN 2 AND 3;")
        assert rows == [(3,)]
        for bad in ["SELECT x FROM T;", "SELECT k FROM NOPE;", "SELECT k FROM T WHERE;", "INSERT INTO T VALUES (1);",
                    "SELECT k FROM T LEFT JOIN T2 ON T.k = T2.k;", "CREATE TABLE T (k INT);"]:
            with pytest.raises(mdb.QueryError):
                if bad.startswith("SELECT"):
                    db.query(bad)
                else:
                    db.execute(bad)


def test_readme_query_large_through_sql_api():
    """2^19 x 2^19 rows through INSERT statements: the SQL-level call reaches the radix join+count path"""
    rng = np.random.default_rng(9)
    n = 1 << 19
    a, b = rng.integers(0, n, n), rng.integers(0, n, n)
    with mdb.Database() as db:
        db.execute("CREATE TABLE A (id_a INT);")
        db.execute("CREATE TABLE B (id_b INT);")
        for name, arr in (("A", a), ("B", b)):
            for i in range(0, n, 8192):
                db.execute("INSERT INTO %s VALUES %s;" % (name, ",".join("(%d)" % v for v in arr[i:i + 8192])))
        names, rows = db.query("SELECT id_a, COUNT(*) FROM A INNER JOIN B ON A.id_a = B.id_b GROUP BY id_a;")
        path, _ = db.last_path()
    assert path == capi.PATH_RADIX_JOINCOUNT
    ca, cb = np.bincount(a, minlength=n), np.bincount(b, minlength=n)
    keys = np.flatnonzero((ca > 0) & (cb > 0))
    want = sorted(zip(keys.tolist(), (ca * cb)[keys].tolist()))
    assert sorted(rows) == want


def test_failing_dml_leaves_table_and_mirror_untouched():
    """a statement the semantic phase rejects must not change a single row: host pages (read by DELETE / UPDATE WHERE) and the
    device mirror (read by SELECT) stay in agreement.  The reference rejects such statements before its executor runs
    (semantic_insert.c check_value_types, semantic_update.c)."""
    with mdb.Database() as db:
        db.execute("CREATE TABLE T (k INT, v DOUBLE);")
        db.execute("INSERT INTO T VALUES (1, 0.5), (2, 1.5), (3, 2.5);")
        assert db.query("SELECT k, v FROM T;")[1] == [(1, 0.5), (2, 1.5), (3, 2.5)]  # the mirror exists now
        # multi-row INSERT whose LAST tuple is bad: nothing is inserted
        with pytest.raises(mdb.QueryError):
            db.execute("INSERT INTO T VALUES (4, 3.5), (5, 4.5), ('x', 1.0);")
        assert db.query("SELECT COUNT(*) FROM T;")[1] == [(3,)]
        # UPDATE with a literal of the wrong type: no row loses its value (the first matching row used to be zeroed)
        with pytest.raises(mdb.QueryError):
            db.execute("UPDATE T SET k = 'x' WHERE k >= 1;")
        assert db.query("SELECT k, v FROM T;")[1] == [(1, 0.5), (2, 1.5), (3, 2.5)]
        # host pages and mirror agree: a DELETE (host-side WHERE) finds what a SELECT (device) shows
        assert db.execute("DELETE FROM T WHERE k = 1;") == 1
        assert db.query("SELECT k, v FROM T;")[1] == [(2, 1.5), (3, 2.5)]
        # a good UPDATE still works and reaches the mirror
        assert db.execute("UPDATE T SET v = 9.0 WHERE k = 3;") == 1
        assert db.query("SELECT k, v FROM T;")[1] == [(2, 1.5), (3, 9.0)]


def test_order_by_having_limit_distinct_through_sql():
    """the clauses the reference parses, validates and then ignores (midorisql.y:180-196,203; executor_select.c:1723) are
    executed here; expected rows from sqlite3 on the same statements (the cursor returns them in ORDER BY order)"""
    import sqlite3
    rng = np.random.default_rng(31)
    n = 4000
    rows = [(int(k), int(v)) for k, v in zip(rng.integers(0, 300, n), rng.integers(-50, 50, n))]
    con = sqlite3.connect(":memory:")
    con.execute("CREATE TABLE T (k INT, v INT)")
    con.executemany("INSERT INTO T VALUES (?, ?)", rows)
    with mdb.Database() as db:
        db.execute("CREATE TABLE T (k INT, v INT);")
        for i in range(0, n, 500):
            db.execute("INSERT INTO T VALUES %s;" % ", ".join("(%d, %d)" % r for r in rows[i:i + 500]))
        cases = [
            "SELECT k, COUNT(*) FROM T GROUP BY k ORDER BY COUNT(*) DESC, k LIMIT 10",
            "SELECT k, SUM(v) FROM T GROUP BY k HAVING SUM(v) > 100 ORDER BY k DESC",
            "SELECT k, COUNT(*), MAX(v) FROM T WHERE v >= 0 GROUP BY k HAVING COUNT(*) >= 8 AND MAX(v) < 49 ORDER BY k LIMIT 5, 20",
            "SELECT DISTINCT k FROM T WHERE v > 45 ORDER BY k",
            "SELECT DISTINCT v FROM T ORDER BY v DESC LIMIT 7",
        ]
        for sql in cases:
            names, got = db.query(sql + ";")
            want = [tuple(r) for r in con.execute(sql.replace("LIMIT 5, 20", "LIMIT 20 OFFSET 5")).fetchall()]
            # result columns come in the reference's scaffold (hashtable) order, not in SELECT order: compare by name
            sel = sql[len("SELECT "):sql.index(" FROM")].replace("DISTINCT ", "").split(", ")
            pos = [names.index(("T." + c) if "(" not in c else c.replace("(v)", "(T.v)")) for c in sel]
            assert [tuple(r[p] for p in pos) for r in got] == want, sql
        with pytest.raises(mdb.QueryError):
            db.query("SELECT k FROM T ORDER BY v;")  # not in the select list
