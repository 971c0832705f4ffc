#!/usr/bin/env python
"""README query on auto-increment ids (sorted key columns): A.id = 0..n-1, B.id = a random permutation or 0..n-1.
    python profiles/bench_sorted_join.py [--log2-rows 28] [--b sorted|uniform]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from midoridb_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-rows", type=int, default=28)
ap.add_argument("--b", default="uniform")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
n = 1 << args.log2_rows
be = capi.Backend(0)
I = capi.CT_INTEGER
ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
ta.generate(n, [capi.GenSpec(kind=capi.GEN_SEQUENCE, lo=0, hi=n - 1, seed=1)])
tb.generate(n, [capi.GenSpec(kind=capi.GEN_SEQUENCE if args.b == "sorted" else capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=2)])
plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)])
for _ in range(3):
    res = be.select(plan); st = be.stats(); groups = res.nrows; res.free()
be.sync(); be.event_record(0)
for _ in range(args.steps):
    res = be.select(plan); res.free()
be.event_record(1); be.sync()
ms = be.event_elapsed_ms(0, 1) / args.steps
print(json.dumps({"workload": "README query, A.id = 0..2^%d-1 (sorted), B %s" % (args.log2_rows, args.b), "path": st.path,
                  "ms_per_step": ms, "rows_per_s": 2 * n / (ms / 1000.0), "groups": groups, "phase_ms": list(st.phase_ms)}))
