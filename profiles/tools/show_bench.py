#!/usr/bin/env python
"""print the interesting parts of a bench.py JSON line:  python profiles/tools/show_bench.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "n_gpus", "verified", "gpu_launches", "result_groups")})
print("roofline", d.get("roofline"))
rs = d.get("roofline_step") or {}
print("step frac", rs.get("frac"), rs.get("phase_ms"))
print("clocks", d.get("clocks"))
print("nvlink", d.get("nvlink"))
e = d.get("e2e") or {}
print("e2e cold", {k: e.get(k) for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "h2d_gbs")})
print("e2e warm", e.get("warm"))
print("query_execute", e.get("query_execute"))
print("cpu_baseline", d.get("cpu_baseline"))
for k, v in (d.get("extra_configs") or {}).items():
    print(k, {x: v.get(x) for x in ("ms_per_query", "rows_per_s", "frac", "kernel_frac", "verified", "path", "groups")}, "cpu:", v.get("cpu_baseline"))
