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
    if rank != 0:
        return
    from oracle import refdb
    n = args.ref_rows
    if refdb.available():
        kind, sample = "reference", "README query, %dx%d rows, unique INT keys, reference nested-loop executor, 1 thread" % (n, n)
        for _ in range(args.warmup):
            reference_step(n, 0)
        t = [reference_step(n, 1 + i) for i in range(args.steps)]
        ms = 1000.0 * float(np.mean(t))
        value = 2 * n / (ms / 1000.0)
    else:
        kind = "port"
        l2 = 22
        sample = "README query, 2^%d x 2^%d rows, uniform keys, oracle hash join+count (C port), 1 thread" % (l2, l2)
        for _ in range(args.warmup):
            port_step(l2)
        t = [port_step(l2)[0] for _ in range(args.steps)]
        ms = 1000.0 * float(np.mean(t))
        value = 2 * (1 << l2) / (ms / 1000.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "README query A INNER JOIN B ON id GROUP BY id COUNT(*), 2^%d x 2^%d rows, INT (int64) keys uniform in [0,2^%d)"
                        % (args.log2_rows, args.log2_rows, args.log2_rows),
            "rows_per_table": 1 << args.log2_rows, "key_domain": 1 << args.log2_rows,
            "parallelism": "1 process per GPU, %d rank(s), key-range partitioned" % world,
            "l2_hygiene": "inputs (%.1f GB) far larger than the 126 MB L2" % (16.0 * (1 << args.log2_rows) / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-rows", type=int, default=28)
    ap.add_argument("--ref-rows", type=int, default=3072, help="rows per table of the reference arm's bounded sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
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
    if st.path != capi.PATH_RADIX_JOINCOUNT:
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
    rows_per_step = 2 * n_total
    value = rows_per_step / (ms_per_step / 1000.0)
    phase /= args.steps
    alg_bytes = 16.0 * n_total + 16.0 * total_groups  # 8|A| + 8|B| + 16 G

    # dominant kernel: k_radix_partition (two launches per step, one per join side); 8 bytes of key per row in
    part_ms_per_launch = dist.max(phase[1]) / 2.0
    part_bytes_per_launch = 8.0 * n_local
    achieved = part_bytes_per_launch / (part_ms_per_launch / 1000.0) / 1e9 if part_ms_per_launch > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_radix_partition_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_radix_partition", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": part_bytes_per_launch, "ms_per_launch": part_ms_per_launch}
    step_gbs = alg_bytes / (ms_per_step / 1000.0) / 1e9
    roofline_step = {"bound": "hbm", "what": "whole step: (8|A| + 8|B| + 16 G) bytes / step device time", "achieved": step_gbs,
                     "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "algorithmic_bytes_per_step": alg_bytes,
                     "phase_ms": {"partition": phase[1], "histogram_join_emit": phase[2], "exchange_push": phase[6], "barriers_and_other": phase[7]}}

    # multi-GPU: bytes this rank pushed to its peers over NVLink per step, against the nominal NVLink 5 rate per direction.
    # Only the push of join side B is exposed (phase "exchange_push"); side A's push overlaps pass 1 of side B.
    nvlink = None
    if world > 1:
        sent = dist.max(float(st.exchange_bytes))
        push_ms = dist.max(phase[6])
        nvlink = {"bytes_pushed_per_rank_per_step": sent, "exposed_push_ms": push_ms,
                  "achieved_gbs_exposed_half": (sent / 2.0) / (push_ms / 1000.0) / 1e9 if push_ms > 0 else None,
                  "peak_gbs_per_direction": 900.0, "peak_source": "nominal NVLink 5 (18 links x 50 GB/s), not measured here",
                  "what": "2-byte remainders of the partitions owned by peers; 8-byte keys never cross the link"}

    # ---- e2e: the same query through the C ABI from HOST page images (reference row format, pinned memory)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, be, dist, ta, tb, n_local, rows_per_step, flags)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(args)

    ta.drop()
    tb.drop()
    be.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "config": workload_config(args, world), "wall_ms_per_step": wall_ms / args.steps,
                "result_groups": int(total_groups), "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "roofline_step": roofline_step, "nvlink": nvlink, "e2e": e2e, "cpu_baseline": cpu_baseline}
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
           "what": "mdbcu_table_create + mdbcu_table_append_pages (host page images, reference row format, pinned) x2, "
                   "mdbcu_select, mdbcu_result_fetch_columns to pinned host memory, per step"}
    for a in (pa, pb, out_keys, out_cnts):
        be.host_free(a)
    return out


def run_cpu_baseline(args):
    from oracle import refdb
    if refdb.available():
        n = args.ref_rows
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
