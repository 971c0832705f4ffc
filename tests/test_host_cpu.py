"""CPU tests of the host side (libmidoridb_b200.so): the library loads and exports the reference's public API,
the result-column (scaffold) order restatement matches what the reference produced for every golden case, and
the SQL front-end emits the reference's token protocol (token scripts verified against the reference: SURVEY.md
appendix A)."""
import ctypes as C

import pytest

from midoridb_b200 import db as mdb
from tests import helpers

CASES = helpers.load_golden()
API = ["database_open", "database_close", "query_execute", "query_cur_step", "query_column_int64", "query_free"]


def test_api_symbols_exported():
    L = mdb.load_library()
    for name in API + ["query_column_double", "query_column_is_null"]:
        assert hasattr(L, name), name


def test_every_function_declared_in_the_public_header_is_exported():
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "midoridb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = re.findall(r"^\s*(?:struct\s+\w+\s*\*|[A-Za-z_][\w ]*?[\s\*])\s*(\w+)\s*\([^;{]*\)\s*;", hdr, flags=re.M)
    assert set(API) <= set(declared)
    L = mdb.load_library()
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match_reference_abi():
    # include/primitive/column.h:30-49, table.h:23-42, datablock.h:9-13, query.h:24-40 on x86-64
    assert C.sizeof(mdb.Column) == 144
    assert mdb.Table.column_count.offset == 128 + 128 * 144
    assert C.sizeof(mdb.Table) == 128 + 128 * 144 + 8 + 8 + 8 + 40
    assert mdb.QueryOutput.results.offset == 8
    assert mdb.QueryOutput.error.offset == 32
    assert mdb.QueryOutput.n_rows_aff.offset == 32 + 1024


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_scaffold_order_matches_reference(case):
    """result columns come out in the reference's hashtable order, not in SELECT-list order"""
    put = (["COUNT(*)"] if "COUNT(*)" in case["columns"] else [])
    for tbl in case["tables"]:
        put += ["%s.%s" % (tbl["name"], c) for c in tbl["cols"]]
    order = mdb.scaffold_order(put)
    assert [k for k in order if k in case["columns"]] == case["columns"]


def _tokens(sql):
    L = C.CDLL(mdb.LIB_PATH)
    out = []
    CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p)
    cb = CB(lambda ctx, tok: out.append(tok.decode()) or 0)
    err = C.create_string_buffer(256)
    L.mdb_sql_to_tokens.argtypes = [C.c_char_p, CB, C.c_void_p, C.c_char_p, C.c_size_t]
    rc = L.mdb_sql_to_tokens(sql.encode(), cb, None, err, 256)
    return rc, out, err.value.decode()


def test_front_end_token_protocol():
    # scripts that executed correctly on the reference
    rc, toks, _ = _tokens("SELECT id_a, COUNT(*) FROM A INNER JOIN B ON A.id_a = B.id_b GROUP BY id_a;")
    assert rc == 0 and toks == ["NAME id_a", "COUNTALL", "TABLE A", "TABLE B", "FIELDNAME A.id_a", "FIELDNAME B.id_b", "CMP 4",
                                "ONEXPR", "JOIN 1", "NAME id_a", "GROUPBYLIST 1", "SELECT 0 4", "STMT"]
    rc, toks, _ = _tokens("SELECT * FROM A JOIN B ON A.id_a = B.id_b JOIN C ON A.id_a = C.id_c;")
    assert toks == ["SELECTALL", "TABLE A", "TABLE B", "FIELDNAME A.id_a", "FIELDNAME B.id_b", "CMP 4", "ONEXPR", "JOIN 1", "TABLE C",
                    "FIELDNAME A.id_a", "FIELDNAME C.id_c", "CMP 4", "ONEXPR", "JOIN 1", "SELECT 0 2", "STMT"]
    rc, toks, _ = _tokens("SELECT k FROM T WHERE k >= 5 AND k <= 9;")
    assert toks == ["NAME k", "TABLE T", "NAME k", "NUMBER 5", "CMP 6", "NAME k", "NUMBER 9", "CMP 5", "AND", "WHERE", "SELECT 0 3", "STMT"]
    # BETWEEN is lowered to the same tokens (the reference has the keyword but no grammar rule)
    rc, toks2, _ = _tokens("SELECT k FROM T WHERE k BETWEEN 5 AND 9;")
    assert rc == 0 and toks2 == toks
    rc, toks, _ = _tokens("SELECT COUNT(*) FROM A WHERE id_a > 1;")
    assert toks == ["COUNTALL", "TABLE A", "NAME id_a", "NUMBER 1", "CMP 2", "WHERE", "SELECT 0 3", "STMT"]
    rc, toks, _ = _tokens("SELECT f1 FROM A WHERE f1 IN (123, 789);")
    assert toks == ["NAME f1", "TABLE A", "NAME f1", "NUMBER 123", "NUMBER 789", "ISIN 2", "WHERE", "SELECT 0 3", "STMT"]
    rc, toks, _ = _tokens("CREATE TABLE A (id_a INT, f1 DOUBLE NOT NULL);")
    assert toks == ["STARTCOL", "COLUMNDEF 50000 id_a", "STARTCOL", "ATTR NOTNULL", "COLUMNDEF 80000 f1", "CREATE 0 2 A", "STMT"]
    rc, toks, _ = _tokens("INSERT INTO A VALUES (1, -12345), (NULL, 2 + 3);")
    assert toks == ["NUMBER 1", "NUMBER -12345", "VALUES 2", "NULL", "NUMBER 2", "NUMBER 3", "ADD", "VALUES 2", "INSERTVALS 0 2 A", "STMT"]
    rc, toks, _ = _tokens("UPDATE A SET f1 = 7 WHERE id_a = 2 OR f1 IS NULL;")
    assert toks == ["NUMBER 7", "ASSIGN f1", "NAME id_a", "NUMBER 2", "CMP 4", "NAME f1", "ISNULL", "OR", "WHERE", "UPDATE A 1 1", "STMT"]
    rc, toks, _ = _tokens("DELETE FROM A WHERE id_a <> 3;")
    assert toks == ["NAME id_a", "NUMBER 3", "CMP 3", "WHERE", "DELETEONE A", "STMT"]
    # operator precedence (midorisql.y:49-63): OR < XOR < AND < comparison < + < *
    rc, toks, _ = _tokens("SELECT a FROM T WHERE a = 1 OR b = 2 AND c < 3 + 4 * 5;")
    assert toks[2:-2] == ["NAME a", "NUMBER 1", "CMP 4", "NAME b", "NUMBER 2", "CMP 4", "NAME c", "NUMBER 3", "NUMBER 4", "NUMBER 5",
                         "MUL", "ADD", "CMP 1", "AND", "OR", "WHERE"]
    for bad in ["SELECT FROM A;", "SELECT a FROM A", "SELEC a FROM A;", "SELECT a FROM A WHERE;", "INSERT INTO A VALUES (1;"]:
        rc, _, msg = _tokens(bad)
        assert rc != 0 and msg
