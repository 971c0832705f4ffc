#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[4]): small-build / large-probe star join.
    D(id, g): 65 536 rows, id = permutation of [0, 65536), g uniform in [0, 1024)
    F(fk, m): 2^log2_rows rows, fk uniform in [0, 65536), m uniform int64
    SELECT g, MIN(m), MAX(m) FROM D INNER JOIN F ON id = fk GROUP BY g          (16 B per fact row are read once)
    python profiles/bench_star_join.py [--log2-rows 30] [--steps 10]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from midoridb_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-rows", type=int, default=30)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
n = 1 << args.log2_rows
be = capi.Backend(0)
I = capi.CT_INTEGER
td, tf = be.create_table("D", [I, I]), be.create_table("F", [I, I])
td.generate(65536, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=65535, seed=5), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=1023, seed=6)])
tf.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=65535, seed=7), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 31) - 1, seed=8)])
plan = capi.make_plan([td, tf], joins=[((0, 0), (1, 0))], group=[(0, 1)],
                      out=[(capi.OUT_COLUMN, 0, 1), (capi.OUT_MIN, 1, 1), (capi.OUT_MAX, 1, 1)])
for _ in range(3):
    res = be.select(plan); st = be.stats(); groups = res.nrows; res.free()
if st.path != capi.PATH_DIRECT_STAR:
    raise SystemExit("the star-join path did not run (path=%d)" % st.path)
be.sync(); be.event_record(0)
for _ in range(args.steps):
    res = be.select(plan); st = be.stats(); res.free()
be.event_record(1); be.sync()
ms = be.event_elapsed_ms(0, 1) / args.steps
peak = 6533.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gbs = 16.0 * n / (ms / 1000.0) / 1e9
print(json.dumps({"workload": "D(65536) JOIN F(2^%d) GROUP BY g MIN(m), MAX(m)" % args.log2_rows, "ms_per_query": ms,
                  "probe_kernel_ms": st.dominant_ms, "rows_per_s": n / (ms / 1000.0), "achieved_gbs": gbs, "peak_gbs": peak,
                  "frac": gbs / peak, "kernel_frac": 16.0 * n / (st.dominant_ms / 1000.0) / 1e9 / peak, "groups": groups}))
