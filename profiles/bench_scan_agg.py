#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[1]): the fused scan + filter + aggregate kernel.

    SELECT COUNT(*), SUM(v) FROM T WHERE k BETWEEN lo AND hi      T(k BIGINT, v DOUBLE), 2^log2_rows rows

Algorithmic bytes = 16 B per row (both columns are read once).  Prints one JSON line; not part of bench.py's contract.
    python profiles/bench_scan_agg.py [--log2-rows 30] [--steps 10] [--warmup 3] [--selectivity 0.5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from midoridb_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-rows", type=int, default=30)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--selectivity", type=float, default=0.5)
    args = ap.parse_args()
    n = 1 << args.log2_rows
    be = capi.Backend(0)
    t = be.create_table("T", [capi.CT_INTEGER, capi.CT_DOUBLE])
    t.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 40) - 1, seed=3),
                   capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, lo=0, hi=1, seed=4)])
    lo, hi = 0, int(args.selectivity * (1 << 40))
    plan = capi.make_plan([t], pred=[("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)],
                          out=[(capi.OUT_COUNT_STAR,), (capi.OUT_SUM, 0, 1)])
    for _ in range(args.warmup):
        res = be.select(plan)
        path = be.stats().path
        res.free()
    if path != capi.PATH_SCAN_AGG:
        raise SystemExit("the fused scan+aggregate path did not run (path=%d)" % path)
    be.sync()
    be.event_record(0)
    for _ in range(args.steps):
        res = be.select(plan)
        cols = res.fetch_columns()
        res.free()
    be.event_record(1)
    be.sync()
    ms = be.event_elapsed_ms(0, 1) / args.steps
    peak = 6533.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = 16.0 * n / (ms / 1000.0) / 1e9
    print(json.dumps({"workload": "SELECT COUNT(*), SUM(v) FROM T WHERE k BETWEEN lo AND hi, 2^%d rows BIGINT/DOUBLE" % args.log2_rows,
                      "ms_per_query": ms, "rows_per_s": n / (ms / 1000.0), "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
                      "count": int(cols[0][0][0]), "selectivity": args.selectivity,
                      "note": "query time includes the result fetch (16 bytes) and the host round trip"}))
    t.drop()
    be.close()


if __name__ == "__main__":
    main()
