#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[3]): three-way join with WHERE and GROUP BY SUM/AVG, Zipf(1.1) foreign keys.
Runs on the GENERAL operators (hash join as CSR multimap, predicate interpreter, hash aggregate); no fused path yet.
    A(id, x DOUBLE in [0,1)): 2^log2_rows rows, id Zipf(1.1) over [0, 2^dim);  B(id, y), C(id, z): 2^dim rows, unique ids
    SELECT A.id, SUM(A.x), AVG(C.z) FROM A JOIN B ON A.id = B.id JOIN C ON A.id = C.id
      WHERE A.x >= 0.25 AND B.y < 500 GROUP BY A.id
    python profiles/bench_threeway.py [--log2-rows 24] [--log2-dim 18] [--steps 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from midoridb_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-rows", type=int, default=24)
ap.add_argument("--log2-dim", type=int, default=18)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
n, nd = 1 << args.log2_rows, 1 << args.log2_dim
be = capi.Backend(0)
I, D = capi.CT_INTEGER, capi.CT_DOUBLE
ta, tb, tc = be.create_table("A", [I, D]), be.create_table("B", [I, I]), be.create_table("C", [I, I])
ta.generate(n, [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=nd - 1, param=1.1, seed=11), capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, seed=12)])
tb.generate(nd, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=nd - 1, seed=13), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=999, seed=14)])
tc.generate(nd, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=nd - 1, seed=15), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=49, seed=16)])
plan = capi.make_plan([ta, tb, tc], joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
                      pred=[("col", 0, 1), ("dbl", 0.25), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
                      group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_SUM, 0, 1), (capi.OUT_AVG, 2, 1)])
for _ in range(4):  # the first queries grow the memory pool (360 ms for the very first one)
    res = be.select(plan); st = be.stats(); groups = res.nrows; res.free()
be.sync(); be.event_record(0)
for _ in range(args.steps):
    res = be.select(plan); st = be.stats(); res.free()
be.event_record(1); be.sync()
ms = be.event_elapsed_ms(0, 1) / args.steps
rows = n + 2 * nd
print(json.dumps({"workload": "A(2^%d, Zipf 1.1) JOIN B JOIN C (2^%d each) WHERE ... GROUP BY A.id SUM, AVG" % (args.log2_rows, args.log2_dim),
                  "path": st.path, "ms_per_query": ms, "rows_per_s": rows / (ms / 1000.0), "groups": groups, "phase_ms": list(st.phase_ms),
                  "kernel_launches": st.kernel_launches}))
