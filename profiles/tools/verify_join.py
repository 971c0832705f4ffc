#!/usr/bin/env python
"""Debug helper: README query at 2^k x 2^k uniform keys, result checked against an independent torch histogram
(bench.verify_join_count) several times in a row.   python profiles/tools/verify_join.py [log2_rows] [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from midoridb_b200 import capi  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
be = capi.Backend(0)
I = capi.CT_INTEGER
ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
ta.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=1)])
tb.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=2)])
plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)])
for i in range(reps):
    res = be.select(plan)
    v = bench.verify_join_count(be, ta, tb, res, n, bench.Dist(0, 1), 0)
    print("run %d: groups %d path %d verified %s" % (i, res.nrows, be.stats().path, v))
    res.free()
