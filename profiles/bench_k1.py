"""K1 of the north star on its own (for ncu launch lists): python profiles/bench_k1.py [log2_rows]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from midoridb_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("log2_rows", nargs="?", type=int, default=28)
ap.add_argument("--no-verify", action="store_true")
a = ap.parse_args()
with capi.Backend(0) as be:
    print(json.dumps(bench.run_k1(be, a, bench.measured_peaks()[0], log2_rows=a.log2_rows)))
