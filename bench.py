#!/usr/bin/env python
"""bench.py - headline benchmark of the MidoriDB hot path on B200 (contract: see the task brief / DESIGN.md).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
    SELECT id_a, COUNT(*) FROM A INNER JOIN B ON A.id_a = B.id_b GROUP BY id_a
    A, B: 2^28 rows each, one INT (int64) key column, keys uniform in [0, 2^28)  (synthetic, generated on device)
A "step" is one execution of that query through the C ABI (mdbcu_select) with both tables resident in HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
    python bench.py --impl reference [...]                          the reference's own CPU executor (oracle/_ref)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "join_groupby_rows_per_s"
UNIT = "rows/s"
QUERY = "SELECT id_a, COUNT(*) FROM A INNER JOIN B ON A.id_a = B.id_b GROUP BY id_a;"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def count(self):
        try:
            self.f.flush()
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax = float(p[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class Dist:
    """rendezvous / barrier / max-over-ranks through torch.distributed (gloo on CPU tensors: the data path's only
    collective is NCCL inside libmidoridb_cuda.so)"""

    def __init__(self, rank, world):
        self.rank, self.world, self.td = rank, world, None
        if world > 1:
            import torch
            import torch.distributed as td
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            td.init_process_group("gloo", rank=rank, world_size=world)
            self.td, self.torch = td, torch

    def barrier(self):
        if self.td:
            self.td.barrier()

    def max(self, v):
        if not self.td:
            return v
        t = self.torch.tensor([float(v)], dtype=self.torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t[0])

    def sum(self, v):
        if not self.td:
            return v
        t = self.torch.tensor([float(v)], dtype=self.torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t[0])

    def sum_cuda(self, t):
        """all-reduce of a CUDA tensor (verification only; NCCL group created on first use)"""
        if not self.td:
            return t
        if not hasattr(self, "nccl"):
            self.nccl = self.td.new_group(backend="nccl")
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM, group=self.nccl)
        return t

    def bcast_bytes(self, b, n):
        if not self.td:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8)
        if self.rank == 0:
            t = self.torch.tensor(list(b), dtype=self.torch.uint8)
        self.td.broadcast(t, 0)
        return bytes(t.tolist())

    def close(self):
        if self.td:
            self.td.destroy_process_group()


def verify_join_count(be, ta, tb, res, key_domain, dist, local_gpu):
    """Full content check of one README-query result at benchmark size, independent of the library's kernels:
    torch.bincount over the key columns (zero-copy views of the mirror), summed over the ranks, gives cntA and cntB;
    every (key, count) row of this rank's result must equal cntA[key] * cntB[key] > 0, keys must be distinct, and the
    number of result rows over all ranks must equal the number of keys present on both sides.
    Matches the cardinality-asserting style of the reference's own test_select_11 (tests/engine/executor_select.c:348-374)."""
    import torch
    from midoridb_b200 import capi
    dev = torch.device("cuda", local_gpu)

    def hist(t):
        ptr, n = t.device_ptr(0)
        if n == 0:
            return torch.zeros(key_domain, dtype=torch.int64, device=dev)
        keys = torch.as_tensor(capi.DeviceArray(ptr, n), device=dev)
        if int(keys.min()) < 0 or int(keys.max()) >= key_domain:
            raise AssertionError("generated keys outside [0, %d)" % key_domain)
        return torch.bincount(keys, minlength=key_domain)

    ca, cb = hist(ta), hist(tb)
    if dist.world > 1:
        ca, cb = dist.sum_cuda(ca), dist.sum_cuda(cb)
    expect = ca * cb
    del ca, cb
    want_groups = int(torch.count_nonzero(expect))
    ok = True
    n = res.nrows
    if n:
        k = torch.as_tensor(capi.DeviceArray(res.device_ptr(0), n), device=dev)
        c = torch.as_tensor(capi.DeviceArray(res.device_ptr(1), n), device=dev)
        ok = ok and int(k.min()) >= 0 and int(k.max()) < key_domain
        ok = ok and bool(torch.equal(expect[k], c)) and int(c.min()) >= 1
        ok = ok and int(torch.unique(k).numel()) == n
    total = int(dist.sum(n))
    ok = ok and total == want_groups
    return bool(dist.sum(0 if ok else 1) == 0)


def reference_step(n_rows, seed):
    """README query on the UNMODIFIED reference executor (oracle/_ref), unique keys (its correct domain at any size)"""
    from oracle import refdb
    rng = np.random.default_rng(seed)
    a = rng.permutation(n_rows).astype(np.int64)
    b = rng.permutation(n_rows).astype(np.int64)
    with refdb.RefDatabase() as db:
        ta = db.create_table("A", ["id_a"], [refdb.CT_INTEGER])
        tb = db.create_table("B", ["id_b"], [refdb.CT_INTEGER])
        db.append(ta, a)
        db.append(tb, b)
        t0 = time.perf_counter()
        res = db.query(QUERY)
        dt = time.perf_counter() - t0
        assert res.cells.shape[0] == n_rows
    return dt


def port_step(log2_rows):
    """the oracle's single-thread hash join+count over raw key arrays (kind 'port')"""
    from oracle import oracle
    n = 1 << log2_rows
    rng = np.random.default_rng(1)
    a = rng.integers(0, n, n).astype(np.int64)
    b = rng.integers(0, n, n).astype(np.int64)
    t0 = time.perf_counter()
    groups = oracle.join_count_groups(a, b)
    dt = time.perf_counter() - t0
    return dt, groups


def run_reference_arm(args, rank, world):
    """the reference's own CPU executor on a BOUNDED SAMPLE of the workload; `config` states what actually ran"""
    if rank != 0:
        return
    from oracle import refdb
    if args.config == 2:
        n = 1 << args.ref_scan_log2
        kind = "reference" if refdb.available() else "port"
        for _ in range(min(args.warmup, 1)):
            cpu_scan_step(n, kind)
        t = [cpu_scan_step(n, kind) for _ in range(args.steps)]
        ms = 1000.0 * float(np.mean(t))
        value, metric = n / (ms / 1000.0), "scan_filter_aggregate_rows_per_s"
        sample = ("SELECT COUNT(*) FROM T WHERE k >= lo AND k <= hi on 2^%d rows (the reference has no SUM; filter + COUNT only), "
                  "reference executor, 1 thread" % args.ref_scan_log2)
        cfg = {"workload": "BOUNDED SAMPLE of config 2: " + sample, "rows_per_table": n,
               "sample_of": "SELECT COUNT(*), SUM(v) FROM T WHERE k BETWEEN lo AND hi, 2^30 rows BIGINT/DOUBLE"}
    else:
        n = args.ref_rows
        if refdb.available():
            kind, sample = "reference", "README query, %d x %d rows, unique INT keys, reference nested-loop executor, 1 thread" % (n, n)
            for _ in range(min(args.warmup, 1)):
                reference_step(n, 0)
            t = [reference_step(n, 1 + i) for i in range(args.steps)]
            ms = 1000.0 * float(np.mean(t))
            value = 2 * n / (ms / 1000.0)
        else:
            kind, n = "port", 1 << 22
            sample = "README query, 2^22 x 2^22 rows, uniform keys, oracle hash join+count (C port), 1 thread"
            for _ in range(min(args.warmup, 1)):
                port_step(22)
            t = [port_step(22)[0] for _ in range(args.steps)]
            ms = 1000.0 * float(np.mean(t))
            value = 2 * n / (ms / 1000.0)
        metric = METRIC
        cfg = {"workload": "BOUNDED SAMPLE of the headline workload: " + sample + " (O(n^2): its rows/s do not carry over to 2^28 rows)",
               "rows_per_table": n, "key_domain": n, "sample_of": workload_config(args, world)["workload"]}
    line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_scan_step(n, kind):
    """filter + COUNT scan on the CPU: the unmodified reference (oracle/_ref) or, without it, the oracle port"""
    rng = np.random.default_rng(3)
    k = rng.integers(0, 1 << 31, n).astype(np.int64)
    lo, hi = 1 << 29, 3 * (1 << 29) - 1
    if kind == "reference":
        from oracle import refdb
        with refdb.RefDatabase() as db:
            t = db.create_table("T", ["k"], [refdb.CT_INTEGER])
            db.append(t, k)
            t0 = time.perf_counter()
            db.query("SELECT COUNT(*) FROM T WHERE k >= %d AND k <= %d;" % (lo, hi))
            return time.perf_counter() - t0
    from midoridb_b200 import capi
    from oracle import oracle
    ot = oracle.OracleTable([capi.CT_INTEGER])
    ot.append_columns([k])
    plan = capi.make_plan([ot], pred=[("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)],
                          out=[(capi.OUT_COUNT_STAR,)])
    t0 = time.perf_counter()
    oracle.select(plan)
    return time.perf_counter() - t0


def workload_config(args, world):
    return {"workload": "README query A INNER JOIN B ON id GROUP BY id COUNT(*), 2^%d x 2^%d rows, INT (int64) keys uniform in [0,2^%d)"
                        % (args.log2_rows, args.log2_rows, args.log2_rows),
            "rows_per_table": 1 << args.log2_rows, "key_domain": 1 << args.log2_rows,
            "parallelism": "1 process per GPU, %d rank(s), key-range partitioned" % world,
            "l2_hygiene": "inputs (%.1f GB) far larger than the 126 MB L2" % (16.0 * (1 << args.log2_rows) / 1e9)}


# ---------------------------------------------------------------------------------------------------------------
# The other configurations of BASELINE.json (2, 4, 5): device-resident inputs, CUDA events on the library's stream,
# every result checked against an independent torch computation over zero-copy views of the mirrored columns.

def _dev_tensor(t, col, dtype_str="<i8"):
    import torch
    from midoridb_b200 import capi
    ptr, n = t.device_ptr(col)
    return torch.as_tensor(capi.DeviceArray(ptr, n, dtype_str), device=torch.device("cuda", t.backend.device))


def _timed(be, step, steps, warmup):
    for _ in range(warmup):
        step()
    be.sync()
    be.event_record(2)
    for _ in range(steps):
        st = step()
    be.event_record(3)
    be.sync()
    return be.event_elapsed_ms(2, 3) / steps, st


def run_config2(be, args, peak, log2_rows=30, steps=5, warmup=3):
    """SELECT COUNT(*), SUM(v) FROM T WHERE k BETWEEN lo AND hi   T(k BIGINT, v DOUBLE), 16 B per row read once"""
    import torch
    from midoridb_b200 import capi
    n = 1 << log2_rows
    t = be.create_table("T", [capi.CT_INTEGER, capi.CT_DOUBLE])
    t.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 31) - 1, seed=3),
                   capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, lo=0, hi=1, seed=4)])
    lo, hi = 1 << 29, 3 * (1 << 29) - 1  # 50 % selectivity
    plan = capi.make_plan([t], pred=[("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)],
                          out=[(capi.OUT_COUNT_STAR,), (capi.OUT_SUM, 0, 1)])
    got = {}

    def step():
        res = be.select(plan)
        st = be.stats()
        cols, _ = res.fetch_columns()
        got["count"], got["sum"] = int(cols[0][0]), float(cols[1][0])
        res.free()
        return st

    ms, st = _timed(be, step, steps, warmup)
    if st.path != capi.PATH_SCAN_AGG:
        raise SystemExit("bench: the fused scan+aggregate path did not run (path=%d)" % st.path)
    verified = None
    if not args.no_verify:
        k, v = _dev_tensor(t, 0), _dev_tensor(t, 1, "<f8")
        m = (k >= lo) & (k <= hi)
        want_count = int(m.sum())
        want_sum = float(torch.where(m, v, torch.zeros((), dtype=v.dtype, device=v.device)).sum())
        verified = bool(want_count == got["count"] and abs(want_sum - got["sum"]) <= 1e-9 * abs(want_sum))
        del k, v, m
        torch.cuda.empty_cache()
    t.drop()
    gbs = 16.0 * n / (ms / 1000.0) / 1e9
    return {"config": 2, "workload": "SELECT COUNT(*), SUM(v) FROM T WHERE k BETWEEN lo AND hi, 2^%d rows BIGINT/DOUBLE, 50%% selectivity" % log2_rows,
            "path": "k_scan_filter_aggregate", "ms_per_query": ms, "rows_per_s": n / (ms / 1000.0), "algorithmic_bytes": 16 * n,
            "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "count": got["count"], "verified": verified,
            "tolerance": "COUNT exact, DOUBLE SUM 1e-9 relative (torch float64 reduction)",
            "note": "query time includes the 16-byte result fetch and the host round trip"}


def run_k1(be, args, peak, log2_rows=28, steps=5, warmup=2):
    """K1 of the north star on the GENERAL operators: SELECT k, v FROM T WHERE k < lo OR k >= hi (an OR: not the fused scan's
    shape) - predicate evaluation into a verdict bitmap, ballot compaction into a selection vector, gather of the projected
    columns.  Algorithmic bytes: 16 B per input row + 16 B per result row."""
    import torch
    from midoridb_b200 import capi
    n = 1 << log2_rows
    t = be.create_table("T", [capi.CT_INTEGER, capi.CT_INTEGER])
    t.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 31) - 1, seed=5),
                   capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=-(1 << 40), hi=1 << 40, seed=6)])
    lo, hi = 107374182, 2040109465  # 5 % below, 5 % above: 10 % of the rows qualify
    plan = capi.make_plan([t], pred=[("col", 0, 0), ("int", lo), ("cmp", 1), ("col", 0, 0), ("int", hi), ("cmp", 6), ("or",)],
                          out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COLUMN, 0, 1)])
    got = {}

    def step(keep=False):
        res = be.select(plan)
        st = be.stats()
        got["n"] = res.nrows
        if keep:
            return st, res
        res.free()
        return st

    ms, st = _timed(be, step, steps, warmup)
    verified = None
    if not args.no_verify:
        _, res = step(keep=True)
        rk = torch.as_tensor(capi.DeviceArray(res.device_ptr(0), res.nrows, "<i8"), device=torch.device("cuda", be.device))
        rv = torch.as_tensor(capi.DeviceArray(res.device_ptr(1), res.nrows, "<i8"), device=torch.device("cuda", be.device))
        k, v = _dev_tensor(t, 0), _dev_tensor(t, 1)
        m = (k < lo) | (k >= hi)
        verified = bool(int(m.sum()) == res.nrows and int(k[m].sum()) == int(rk.sum()) and int(v[m].sum()) == int(rv.sum())
                        and int((k[m] ^ v[m]).sum()) == int((rk ^ rv).sum()))  # (the pairing of the two columns survives)
        del rk, rv, k, v, m
        res.free()
        torch.cuda.empty_cache()
    t.drop()
    alg = 16 * n + 16 * got["n"]
    gbs = alg / (ms / 1000.0) / 1e9
    return {"config": "K1", "workload": "SELECT k, v FROM T WHERE k < lo OR k >= hi, 2^%d rows BIGINT/BIGINT, 10%% selectivity (general "
                                        "operators: predicate program -> verdict bitmap -> compaction -> gather)" % log2_rows,
            "path": "general operators (k_eval_pred, bitmap compaction, k_gather_out)", "ms_per_query": ms, "rows_per_s": n / (ms / 1000.0),
            "result_rows": int(got["n"]), "algorithmic_bytes": alg, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
            "kernel_launches": int(st.kernel_launches), "phase_ms": [float(x) for x in st.phase_ms], "verified": verified,
            "verified_how": "row count, sums of both columns and the sum of k XOR v against torch on the mirrored columns"}


def run_config5(be, args, peak, log2_rows=30, steps=5, warmup=3):
    """D(id, g) 65 536 rows JOIN F(fk, m) 2^30 rows ON id = fk GROUP BY g MIN(m), MAX(m): 16 B per fact row read once"""
    import torch
    from midoridb_b200 import capi
    n = 1 << log2_rows
    I = capi.CT_INTEGER
    td, tf = be.create_table("D", [I, I]), be.create_table("F", [I, I])
    td.generate(65536, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=65535, seed=5), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=1023, seed=6)])
    tf.generate(n, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=65535, seed=7), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 31) - 1, seed=8)])
    plan = capi.make_plan([td, tf], joins=[((0, 0), (1, 0))], group=[(0, 1)],
                          out=[(capi.OUT_COLUMN, 0, 1), (capi.OUT_MIN, 1, 1), (capi.OUT_MAX, 1, 1)])
    got = {}

    def step():
        res = be.select(plan)
        st = be.stats()
        got["cols"], _ = res.fetch_columns()
        res.free()
        return st

    ms, st = _timed(be, step, steps, warmup)
    if st.path != capi.PATH_DIRECT_STAR:
        raise SystemExit("bench: the star-join path did not run (path=%d)" % st.path)
    verified = None
    if not args.no_verify:
        did, dg, fk, m = _dev_tensor(td, 0), _dev_tensor(td, 1), _dev_tensor(tf, 0), _dev_tensor(tf, 1)
        gmap = torch.empty(65536, dtype=torch.int64, device=did.device)
        gmap[did] = dg
        grp = gmap[fk]
        big = torch.iinfo(torch.int64)
        mn = torch.full((1024,), big.max, dtype=torch.int64, device=did.device).scatter_reduce_(0, grp, m, "amin")
        mx = torch.full((1024,), big.min, dtype=torch.int64, device=did.device).scatter_reduce_(0, grp, m, "amax")
        g_got = torch.from_numpy(np.ascontiguousarray(got["cols"][0])).to(did.device)
        verified = bool(len(got["cols"][0]) == int((mn != big.max).sum()) and
                        torch.equal(mn[g_got].cpu(), torch.from_numpy(np.ascontiguousarray(got["cols"][1]))) and
                        torch.equal(mx[g_got].cpu(), torch.from_numpy(np.ascontiguousarray(got["cols"][2]))))
        del did, dg, fk, m, gmap, grp
        torch.cuda.empty_cache()
    groups = len(got["cols"][0])
    td.drop()
    tf.drop()
    gbs = 16.0 * n / (ms / 1000.0) / 1e9
    return {"config": 5, "workload": "D(65536 rows) JOIN F(2^%d rows) ON id = fk GROUP BY g MIN(m), MAX(m)" % log2_rows,
            "path": "k_star_probe", "ms_per_query": ms, "probe_kernel_ms": st.dominant_ms, "rows_per_s": n / (ms / 1000.0),
            "algorithmic_bytes": 16 * n, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
            "kernel_frac": 16.0 * n / (st.dominant_ms / 1000.0) / 1e9 / peak if st.dominant_ms else None,
            "groups": groups, "verified": verified, "tolerance": "bit-exact (integer MIN/MAX)"}


def run_config4(be, args, peak, log2_rows=26, log2_dim=20, steps=5, warmup=4):
    """A(id Zipf 1.1, x) JOIN B(id, y) JOIN C(id, z) WHERE A.x >= 0.25 AND B.y < 500 GROUP BY A.id SUM(A.x), AVG(C.z)"""
    import torch
    from midoridb_b200 import capi
    n, nd = 1 << log2_rows, 1 << log2_dim
    I, D = capi.CT_INTEGER, capi.CT_DOUBLE
    ta, tb, tc = be.create_table("A", [I, D]), be.create_table("B", [I, I]), be.create_table("C", [I, I])
    ta.generate(n, [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=nd - 1, param=1.1, seed=11), capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, seed=12)])
    tb.generate(nd, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=nd - 1, seed=13), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=999, seed=14)])
    tc.generate(nd, [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=nd - 1, seed=15), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=49, seed=16)])
    plan = capi.make_plan([ta, tb, tc], joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
                          pred=[("col", 0, 1), ("dbl", 0.25), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
                          group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_SUM, 0, 1), (capi.OUT_AVG, 2, 1)])
    got = {}

    def step():
        res = be.select(plan)
        st = be.stats()
        got["n"] = res.nrows
        return st, res

    def timed_step():
        st, res = step()
        res.free()
        return st

    ms, st = _timed(be, timed_step, steps, warmup)
    verified = None
    if not args.no_verify:
        st2, res = step()
        cols, _ = res.fetch_columns()
        res.free()
        aid, ax = _dev_tensor(ta, 0), _dev_tensor(ta, 1, "<f8")
        ymap = torch.empty(nd, dtype=torch.int64, device=aid.device)
        zmap = torch.empty(nd, dtype=torch.int64, device=aid.device)
        ymap[_dev_tensor(tb, 0)] = _dev_tensor(tb, 1)
        zmap[_dev_tensor(tc, 0)] = _dev_tensor(tc, 1)
        m = (ax >= 0.25) & (ymap[aid] < 500)
        ids = aid[m]
        sx = torch.zeros(nd, dtype=torch.float64, device=aid.device).index_add_(0, ids, ax[m])
        sz = torch.zeros(nd, dtype=torch.float64, device=aid.device).index_add_(0, ids, zmap[ids].to(torch.float64))
        cn = torch.bincount(ids, minlength=nd)
        g = torch.from_numpy(np.ascontiguousarray(cols[0])).to(aid.device)
        got_sum = torch.from_numpy(np.ascontiguousarray(cols[1])).to(aid.device)
        got_avg = torch.from_numpy(np.ascontiguousarray(cols[2])).to(aid.device)
        ok = int((cn > 0).sum()) == len(cols[0]) and int(torch.unique(g).numel()) == len(cols[0])
        ok = ok and bool(((got_sum - sx[g]).abs() <= 1e-9 * sx[g].abs()).all())
        want_avg = sz[g] / cn[g].to(torch.float64)
        ok = ok and bool(((got_avg - want_avg).abs() <= 1e-9 * want_avg.abs().clamp_min(1e-300)).all())
        verified = bool(ok)
        del aid, ax, ymap, zmap, m, ids, sx, sz, cn
        torch.cuda.empty_cache()
    for t in (ta, tb, tc):
        t.drop()
    rows = n + 2 * nd
    alg = 16 * n + 2 * 16 * nd + 24 * got["n"]
    gbs = alg / (ms / 1000.0) / 1e9
    names = {capi.PATH_GENERAL: "general operators", capi.PATH_FUSED_MULTIWAY: "k_group_update<StarSource> (fused multiway)"}
    return {"config": 4, "workload": "A(2^%d, Zipf 1.1) JOIN B JOIN C (2^%d each) WHERE A.x >= 0.25 AND B.y < 500 GROUP BY A.id SUM(A.x), AVG(C.z)"
                                    % (log2_rows, log2_dim),
            "path": names.get(st.path, str(st.path)), "ms_per_query": ms, "rows_per_s": rows / (ms / 1000.0), "algorithmic_bytes": alg,
            "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "groups": got["n"], "kernel_launches": int(st.kernel_launches),
            "phase_ms": list(st.phase_ms), "verified": verified, "tolerance": "keys exact, DOUBLE SUM / AVG 1e-9 relative"}


def cpu_port_baseline(config):
    """labelled oracle 'port' baselines for the configurations the reference cannot run at all (no SUM/MIN/MAX/AVG, broken
    second join: SURVEY.md D3): the plain-C restatement, one thread, reduced scale"""
    from midoridb_b200 import capi
    from oracle import oracle
    rng = np.random.default_rng(config)
    I, D = capi.CT_INTEGER, capi.CT_DOUBLE
    if config == 4:
        n, nd = 1 << 20, 1 << 14
        w = 1.0 / np.arange(1, nd + 1) ** 1.1
        a_id = rng.choice(nd, size=n, p=w / w.sum()).astype(np.int64)
        oa, ob, oc = oracle.OracleTable([I, D]), oracle.OracleTable([I, I]), oracle.OracleTable([I, I])
        oa.append_columns([a_id, rng.random(n)])
        ob.append_columns([rng.permutation(nd).astype(np.int64), rng.integers(0, 1000, nd).astype(np.int64)])
        oc.append_columns([rng.permutation(nd).astype(np.int64), rng.integers(0, 50, nd).astype(np.int64)])
        plan = capi.make_plan([oa, ob, oc], joins=[((0, 0), (1, 0)), ((0, 0), (2, 0))],
                              pred=[("col", 0, 1), ("dbl", 0.25), ("cmp", 6), ("col", 1, 1), ("int", 500), ("cmp", 1), ("and",)],
                              group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_SUM, 0, 1), (capi.OUT_AVG, 2, 1)])
        rows, sample = n + 2 * nd, "config 4 at A = 2^20 rows, B = C = 2^14 rows"
    else:
        n = 1 << 22
        od, of = oracle.OracleTable([I, I]), oracle.OracleTable([I, I])
        od.append_columns([rng.permutation(65536).astype(np.int64), rng.integers(0, 1024, 65536).astype(np.int64)])
        of.append_columns([rng.integers(0, 65536, n).astype(np.int64), rng.integers(0, 1 << 31, n).astype(np.int64)])
        plan = capi.make_plan([od, of], joins=[((0, 0), (1, 0))], group=[(0, 1)],
                              out=[(capi.OUT_COLUMN, 0, 1), (capi.OUT_MIN, 1, 1), (capi.OUT_MAX, 1, 1)])
        rows, sample = n, "config 5 at F = 2^22 rows, D = 65536 rows"
    t0 = time.perf_counter()
    oracle.select(plan)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": sample + ", oracle/mdb_oracle.c (plain-C restatement with hash join / hash group-by, NOT the reference: "
                               "the reference cannot execute this query), 1 thread"}


def cpu_scan_baseline(args):
    from oracle import refdb
    kind = "reference" if refdb.available() else "port"
    n = 1 << args.ref_scan_log2
    dt = cpu_scan_step(n, kind)
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
            "sample": "SELECT COUNT(*) FROM T WHERE k >= lo AND k <= hi on 2^%d rows (filter + COUNT only: the reference has no SUM), "
                      "1 of the host's cores" % args.ref_scan_log2}


EXTRA = {2: run_config2, 4: run_config4, 5: run_config5}


def run_single_config(args, rank, world, local):
    """python bench.py --config {2,4,5}: that configuration alone, same JSON contract (single GPU)"""
    from midoridb_b200 import capi
    if rank != 0:
        return
    be = capi.Backend(local)
    peak, peak_src = measured_peaks()
    sampler = ClockSampler(local)
    r = EXTRA[args.config](be, args, peak, steps=args.steps, warmup=args.warmup)
    clocks = sampler.stop()
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_scan_baseline(args) if args.config == 2 else cpu_port_baseline(args.config)
    be.close()
    metric = {2: "scan_filter_aggregate_rows_per_s", 4: "threeway_join_groupby_rows_per_s", 5: "star_join_groupby_rows_per_s"}[args.config]
    line = {"metric": metric, "value": r["rows_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_query"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.config != 5 else "int64", "data": "synthetic", "config": {"workload": r["workload"]},
            "verified": r["verified"], "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": r["path"], "achieved": r["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": r["frac"],
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": r["algorithmic_bytes"]},
            "cpu_baseline": cpu, "detail": r}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-rows", type=int, default=28)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (3 = headline)")
    ap.add_argument("--ref-rows", type=int, default=3072, help="rows per table of the reference arm's bounded sample")
    ap.add_argument("--cpu-rows", type=int, default=10000, help="rows per table of the cpu_baseline leg (BASELINE configs[0]: 10k x 10k)")
    ap.add_argument("--ref-scan-log2", type=int, default=19, help="log2 rows of the reference's filter scan (config 2 baseline)")
    ap.add_argument("--qe-log2-rows", type=int, default=20, help="rows per table of the query_execute leg (SQL INSERTs)")
    ap.add_argument("--no-verify", action="store_true", help="skip the torch cross-check of the results")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs block (configs 2, 4, 5)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank, world, local = dist_env()
    if args.config == 1:
        args.log2_rows = 14  # BASELINE configs[0] (10k x 10k on the CPU reference) at the nearest power of two: a latency configuration
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.config in EXTRA:
        run_single_config(args, rank, world, local)
        return

    from midoridb_b200 import capi
    dist = Dist(rank, world)
    be = capi.Backend(local)
    peak, peak_src = measured_peaks()

    n_total = 1 << args.log2_rows
    n_local = n_total // world
    flags = 0
    if world > 1:
        uid = be.comm_unique_id() if rank == 0 else bytes(128)
        uid = dist.bcast_bytes(uid, 128)
        be.comm_init(rank, world, uid)
        flags = capi.PLAN_DISTRIBUTED

    I = capi.CT_INTEGER
    ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
    # strong scaling: rank r holds rows [r*n/P, (r+1)*n/P) of the same global tables (counter-based generator)
    ta.generate(n_local, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n_total - 1, seed=1)], row_offset=rank * n_local)
    tb.generate(n_local, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n_total - 1, seed=2)], row_offset=rank * n_local)
    if world > 1:
        ta.sync_stats()  # collective: global key range, so every rank partitions identically
        tb.sync_stats()
    plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                          out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)], flags=flags)

    def step():
        res = be.select(plan)
        st = be.stats()
        rows = res.nrows
        res.free()
        return st, rows

    for _ in range(args.warmup):
        st, groups = step()
    if st.path != capi.PATH_RADIX_JOINCOUNT and args.log2_rows >= 20:
        raise SystemExit("bench: the radix join+count path did not run (path=%d)" % st.path)

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = be.stats().total_kernel_launches
    phase = np.zeros(8)
    dist.barrier()
    be.sync()
    be.event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, groups = step()
        phase += np.array(list(st.phase_ms))
    be.event_record(1)
    be.sync()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    dev_ms = be.event_elapsed_ms(0, 1)
    dist.barrier()
    launches = be.stats().total_kernel_launches - launches0
    ms_per_step = dist.max(dev_ms) / args.steps
    # nvidia-smi needs ~100 ms per sample: when the timed region is shorter than that, keep the SAME load running
    # (untimed, same step count on every rank) so that the clocks line is sampled under load
    n_extra = int(min(4000, max(0.0, np.ceil((800.0 - ms_per_step * args.steps) / ms_per_step))))
    for _ in range(n_extra):
        step()
    be.sync()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "timed region + %d further untimed steps of the same load" % n_extra
    total_groups = dist.sum(groups)
    # ---- correctness where the numbers are: one more (untimed) execution, checked row by row against torch histograms
    verified = None
    if not args.no_verify:
        res = be.select(plan)
        verified = verify_join_count(be, ta, tb, res, n_total, dist, local) and res.nrows == groups
        res.free()
        import torch
        torch.cuda.empty_cache()
    rows_per_step = 2 * n_total
    value = rows_per_step / (ms_per_step / 1000.0)
    phase /= args.steps
    alg_bytes = 16.0 * n_total + 16.0 * total_groups  # 8|A| + 8|B| + 16 G

    # dominant kernel: k_radix_partition (two launches per step, one per join side); 8 bytes of key per row in
    part_ms_per_launch = dist.max(phase[1]) / 2.0
    part_bytes_per_launch = 8.0 * n_local
    achieved = part_bytes_per_launch / (part_ms_per_launch / 1000.0) / 1e9 if part_ms_per_launch > 0 else 0.0
    traffic, traffic_src = ncu_traffic(n_local)
    roofline = {"bound": "hbm", "kernel": "k_radix_partition", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": part_bytes_per_launch, "ms_per_launch": part_ms_per_launch}
    step_gbs = alg_bytes / (ms_per_step / 1000.0) / 1e9
    roofline_step = {"bound": "hbm", "what": "whole step: (8|A| + 8|B| + 16 G) bytes / step device time", "achieved": step_gbs,
                     "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "algorithmic_bytes_per_step": alg_bytes,
                     "phase_ms": {"partition": phase[1], "histogram_join_emit": phase[2], "barriers_and_rank_skew": phase[7]}}

    # multi-GPU: bytes of 2-byte remainders that cross NVLink per rank and step.  Neither transfer has a phase of its own: side A is
    # pushed into the owners' arenas by a copy kernel on a few SMs WHILE pass 1 of side B runs, side B is read out of the peers'
    # arenas by pass 2 itself, so pass 2's time bounds the transfer of side B (half of the bytes) from above.
    nvlink = None
    if world > 1:
        moved = dist.max(float(st.exchange_bytes))
        p2_ms = dist.max(phase[2])
        gbs = (moved / 2.0) / (p2_ms / 1000.0) / 1e9 if p2_ms > 0 else None
        nvlink = {"bytes_per_rank_per_step": moved, "pass2_ms": p2_ms, "side_b_pull_gbs_lower_bound": gbs,
                  "peak_gbs_per_direction": 770.0, "frac_lower_bound": gbs / 770.0 if gbs else None,
                  "peak_source": "B200_PROFILING.md: peer copy measured on this pool, 770 GB/s per direction (900 nominal)",
                  "what": "2-byte remainders of the partitions owned by another rank (8-byte keys never cross the link): side A pushed "
                          "during pass 1 of side B (hidden), side B pulled by pass 2 while it counts and emits (the rate is a lower bound)"}

    # ---- e2e: the same query through the C ABI from HOST page images (reference row format, pinned memory)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, be, dist, ta, tb, n_local, rows_per_step, flags)

    ta.drop()
    tb.drop()
    if e2e is not None and rank == 0 and world == 1:
        e2e["query_execute"] = run_query_execute(args)

    # ---- the other BASELINE configurations at full size (single GPU; they do not fit next to the headline tables at N > 1 shards)
    extra = None
    if rank == 0 and world == 1 and not args.no_extra and args.config == 3:
        extra = {}
        for c in (2, 4, 5):
            extra["config_%d" % c] = EXTRA[c](be, args, peak)
        extra["k1_filter_projection"] = run_k1(be, args, peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(args)
        if extra is not None:
            extra["config_2"]["cpu_baseline"] = cpu_scan_baseline(args)
            extra["config_4"]["cpu_baseline"] = cpu_port_baseline(4)
            extra["config_5"]["cpu_baseline"] = cpu_port_baseline(5)
    be.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "config": workload_config(args, world), "wall_ms_per_step": wall_ms / args.steps,
                "result_groups": int(total_groups), "verified": verified,
                "verified_how": "every (key, count) row of every rank == torch.bincount(A) * torch.bincount(B), keys distinct, "
                                "total rows == keys present on both sides",
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "roofline_step": roofline_step, "nvlink": nvlink, "e2e": e2e, "cpu_baseline": cpu_baseline, "extra_configs": extra}
        print(json.dumps(line))
    dist.close()


def run_e2e(args, be, dist, ta, tb, n_local, rows_per_step, flags):
    """per step: H2D of both tables' page images (reference row format) + unpack + query + D2H of the result columns"""
    from midoridb_b200 import capi
    I = capi.CT_INTEGER

    def table_pages(t):
        res = be.select(capi.make_plan([t], out=[(capi.OUT_COLUMN, 0, 0)], flags=capi.PLAN_NO_FASTPATH))
        n = res.page_count()
        host = be.host_array(n * capi.PAGE_SIZE)
        res.fetch_pages(out=host)
        res.free()
        return host, n

    pa, npa = table_pages(ta)
    pb, npb = table_pages(tb)
    cap = n_local if dist.world == 1 else 2 * n_local
    out_keys = be.host_array(8 * cap)
    out_cnts = be.host_array(8 * cap)

    def step():
        ea, eb = be.create_table("A", [I]), be.create_table("B", [I])
        ea.append_pages(pa)
        eb.append_pages(pb)
        if dist.world > 1:
            ea.sync_stats()
            eb.sync_stats()
        res = be.select(capi.make_plan([ea, eb], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                                       out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)], flags=flags))
        groups = res.nrows
        assert groups <= cap
        res.fetch_columns_into([out_keys.ctypes.data, out_cnts.ctypes.data])
        res.free()
        ea.drop()
        eb.drop()
        return groups

    step()  # warm-up (allocations, first-touch)
    dist.barrier()
    be.sync()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        groups = step()
    be.sync()
    ms = 1000.0 * (time.perf_counter() - t0) / args.e2e_steps
    ms = dist.max(ms)
    # spot-check the host-side result of the last step
    k = out_keys.view(np.int64)[:groups]
    c = out_cnts.view(np.int64)[:groups]
    assert groups > 0 and k.min() >= 0 and c.min() >= 1
    out = {"value": rows_per_step / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms,
           "h2d_bytes_per_step": int((npa + npb) * capi.PAGE_SIZE), "d2h_bytes_per_step": int(16 * groups),
           "what": "COLD mirror: mdbcu_table_create + mdbcu_table_append_pages (host page images, reference row format, pinned) x2, "
                   "mdbcu_select, mdbcu_result_fetch_columns to pinned host memory, per step",
           "h2d_gbs": (npa + npb) * capi.PAGE_SIZE / (ms / 1000.0) / 1e9,
           "note": "PCIe-bound: the reference row format carries 32 bytes per 8-byte key"}
    for a in (pa, pb, out_keys, out_cnts):
        be.host_free(a)

    # ---- warm mirror: what a SELECT costs once the tables are mirrored (the design's normal state: INSERT/UPDATE/DELETE keep the
    # mirror in sync page by page): mdbcu_select + the result as page images in the reference's row format on the host
    plan = capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)],
                          out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)], flags=flags)
    res = be.select(plan)
    n_res_pages = res.page_count()
    res.free()
    host_pages = be.host_array((n_res_pages + 1024) * capi.PAGE_SIZE)

    def warm_step():
        res = be.select(plan)
        npg = res.page_count()
        res.fetch_pages(out=host_pages)
        res.free()
        return npg

    warm_step()
    dist.barrier()
    be.sync()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        npg = warm_step()
    be.sync()
    wms = dist.max(1000.0 * (time.perf_counter() - t0) / args.e2e_steps)
    be.host_free(host_pages)
    out["warm"] = {"value": rows_per_step / (wms / 1000.0), "unit": UNIT, "ms_per_step": wms, "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": int(npg * capi.PAGE_SIZE),
                   "what": "WARM mirror (tables already mirrored in HBM): mdbcu_select + mdbcu_result_fetch_pages (result rows as "
                           "page images in the reference row format, what query_cur_step walks) to pinned host memory, per step"}
    return out


def run_query_execute(args):
    """the reference's own public surface (database_open / query_execute / query_cur_step / query_column_int64 / query_free,
    src/engine/query.c:35-146) served by libmidoridb_b200.so: rows INSERTed through SQL, the README query, the cursor walked"""
    import ctypes as C
    from midoridb_b200 import db as mdb
    n = 1 << args.qe_log2_rows
    rng = np.random.default_rng(5)
    with mdb.Database() as d:
        d.execute("CREATE TABLE A (id_a INT);")
        d.execute("CREATE TABLE B (id_b INT);")
        t0 = time.perf_counter()
        for name, seed_keys in (("A", rng.integers(0, n, n)), ("B", rng.integers(0, n, n))):
            for i in range(0, n, 8192):
                d.execute("INSERT INTO %s VALUES %s;" % (name, ",".join("(%d)" % v for v in seed_keys[i:i + 8192])))
        insert_s = time.perf_counter() - t0
        L = d.L
        times = []
        for rep in range(3):
            t0 = time.perf_counter()
            out = L.query_execute(C.byref(d.db), QUERY.encode())
            t1 = time.perf_counter()
            if out.contents.status != mdb.ST_OK_WITH_RESULTS:
                raise SystemExit("query_execute failed: %s" % out.contents.error.decode(errors="replace"))
            rs = C.byref(out.contents.results)
            rows, total = 0, 0
            while L.query_cur_step(rs) == mdb.MIDORIDB_ROW:
                rows += 1
                total += L.query_column_int64(rs, 1) if rep == 2 else 0
            t2 = time.perf_counter()
            L.query_free(out)
            times.append((t1 - t0, t2 - t1))
        path, _ = d.last_path()
    qe_ms = 1000.0 * min(t[0] for t in times)
    return {"value": 2 * n / (qe_ms / 1000.0), "unit": UNIT, "ms_per_query_execute": qe_ms, "rows_per_table": n, "result_rows": rows,
            "cursor_walk_ms_python_ctypes": 1000.0 * times[-1][1], "sql_insert_s": insert_s, "path": int(path),
            "what": "query_execute(db, README query) on tables filled by SQL INSERT through libmidoridb_b200.so (host pages -> device mirror "
                    "kept in sync -> mdbcu_select -> result pages); the cursor walk is timed separately (one ctypes call per row and "
                    "column from Python: FFI overhead, not library time)"}


def ncu_traffic(n_local):
    """DRAM bytes per k_radix_partition launch from the newest `ncu --set full` capture under profiles/ - only when that capture
    was taken on THIS version of the kernel (sha of the pass-1 source) and on the same shard size; else None (stale)"""
    import glob
    import hashlib
    src = os.path.join(ROOT, "midoridb_b200", "csrc", "mdb_radix_pass1.cuh")
    sig = hashlib.sha1(open(src, "rb").read()).hexdigest()[:16]
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")), reverse=True):
        try:
            j = json.load(open(path))
        except Exception:
            continue
        if j.get("pass1_source_sha1_16") == sig and int(j.get("keys_per_launch", 0)) == int(n_local):
            return j.get("k_radix_partition_dram_bytes_per_launch"), os.path.relpath(path, ROOT)
    return None, "no ncu capture of this kernel version at this shard size under profiles/ (sha %s)" % sig


def run_cpu_baseline(args):
    from oracle import refdb
    if refdb.available():
        n = args.cpu_rows
        reference_step(256, 0)
        dt = reference_step(n, 1)
        return {"value": 2 * n / dt, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": dt,
                "sample": "README query on the reference's nested-loop executor (compiled in place, oracle/_ref), %dx%d rows, "
                          "unique keys, 1 of the host's cores; O(n^2): do not extrapolate" % (n, n)}
    dt, _ = port_step(22)
    return {"value": 2 * (1 << 22) / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": "oracle hash join+count (C port, single thread), 2^22 x 2^22 rows, uniform keys"}


if __name__ == "__main__":
    main()
