"""GPU tests (-m gpu) of the multi-GPU code path.

* world of one: a distributed plan's decisions with nobody to exchange with;
* LOOP-BACK: W ranks = W contexts of this process on ONE device (mdbcu_comm_init_local), one host thread
  per rank - k_radix_ship, k_arena_barrier, the multi-source pass 2 and the owner reduction of the direct-count path are
  the kernels a real multi-GPU run uses, so the single-GPU box exercises them;
* the same with one context per device when the box has W GPUs (peer stores over NVLink), and the one-process-per-GPU
  NCCL / CUDA-IPC variant under torchrun (tests/dist_check.py); both skip themselves on smaller boxes.
Every case is checked against the CPU oracle on the whole (unsharded) tables."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

from midoridb_b200 import capi
from midoridb_b200.capi import CT_DOUBLE, CT_INTEGER, OUT_COLUMN, OUT_COUNT_STAR, OUT_MAX, OUT_MIN, OUT_SUM, PLAN_DISTRIBUTED
from oracle import oracle
from tests import helpers

pytestmark = pytest.mark.gpu
I, D = CT_INTEGER, CT_DOUBLE
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JOIN = dict(joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])


def gpu_count():
    import torch
    return torch.cuda.device_count()


def run_ranks(devices, body):
    """body(rank, backend) on one thread per rank; the backends form one in-process communicator.  Returns the list of results."""
    bes = [capi.Backend(d) for d in devices]
    capi.comm_init_local(bes)
    out, err = [None] * len(bes), [None] * len(bes)

    def work(r):
        try:
            out[r] = body(r, bes[r])
        except BaseException as e:  # noqa: BLE001 - reported below, per rank
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(bes))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    alive = [t.is_alive() for t in ts]
    for b in bes:
        if not any(alive):
            b.close()
    assert not any(alive), "a rank hung"
    failed = ["rank %d: %r" % (r, e) for r, e in enumerate(err) if e is not None]
    assert not failed, "; ".join(failed)
    return out


def sharded_join(devices, specs_a, specs_b, n_total, check_path=None):
    """README query on tables generated shard by shard (rank r: rows [r n / W, (r + 1) n / W)); returns per-rank
    (keys, counts, stats path, exchange bytes, column A, valid A, column B, valid B)"""
    W = len(devices)
    n_local = n_total // W

    def body(rank, be):
        ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
        ta.generate(n_local, specs_a, row_offset=rank * n_local)
        tb.generate(n_local, specs_b, row_offset=rank * n_local)
        ta.sync_stats()
        tb.sync_stats()
        res = be.select(capi.make_plan([ta, tb], flags=PLAN_DISTRIBUTED, **JOIN))
        st = be.stats()
        (keys, cnts), _ = res.fetch_columns()
        res.free()
        a, av = ta.read_column(0)
        b, bv = tb.read_column(0)
        ta.drop()
        tb.drop()
        return keys, cnts, st.path, st.exchange_bytes, a, av, b, bv

    return run_ranks(devices, body)


def check_against_oracle(parts, histogram=False):
    """histogram: the expected result from numpy histograms instead of the oracle, whose hash join visits every joined pair
    (3 * 10^10 of them when both sides are Zipf-distributed)"""
    ga = np.concatenate([p[4] for p in parts])
    gav = np.concatenate([p[5] for p in parts])
    gb = np.concatenate([p[6] for p in parts])
    gbv = np.concatenate([p[7] for p in parts])
    if histogram:
        lo = int(min(ga.min(), gb.min()))
        size = int(max(ga.max(), gb.max())) - lo + 1
        prod = np.bincount(ga[gav != 0] - lo, minlength=size) * np.bincount(gb[gbv != 0] - lo, minlength=size)
        nz = np.nonzero(prod)[0]
        want = sorted(zip((nz + lo).tolist(), prod[nz].tolist()))
    else:
        oa, ob = oracle.OracleTable([I]), oracle.OracleTable([I])
        oa.append_columns([ga], [(gav == 0).astype(np.uint8)])
        ob.append_columns([gb], [(gbv == 0).astype(np.uint8)])
        _, cells, _ = oracle.select(capi.make_plan([oa, ob], **JOIN))
        want = sorted(zip(cells[0].tolist(), cells[1].tolist()))
    got = sorted(zip(np.concatenate([p[0] for p in parts]).tolist(), np.concatenate([p[1] for p in parts]).tolist()))
    assert got == want, "distributed result differs from the oracle (%d vs %d groups)" % (len(got), len(want))
    # the ranks own disjoint, ascending key ranges: the per-rank results concatenate without a merge
    last = None
    for p in parts:
        if p[0].size:
            assert last is None or last < p[0].min()
            last = p[0].max()
    return len(want)


def test_distributed_world_of_one_matches_oracle():
    rng = np.random.default_rng(77)
    na, nb = 700001, 650000
    a = rng.integers(-1000, 1 << 20, na)
    b = rng.integers(500, (1 << 20) + 7000, nb)
    an = (rng.random(na) < 0.02).astype(np.uint8)
    with capi.Backend(0) as be:
        be.comm_init(0, 1, be.comm_unique_id())
        ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
        ta.append_columns([a], [an])
        tb.append_columns([b])
        with pytest.raises(capi.MdbError):  # global statistics are mandatory for distributed plans
            be.select(capi.make_plan([ta, tb], flags=PLAN_DISTRIBUTED, **JOIN))
        ta.sync_stats()
        tb.sync_stats()
        res = be.select(capi.make_plan([ta, tb], flags=PLAN_DISTRIBUTED, **JOIN))
        st = be.stats()
        got = res.rows()
        res.free()
        assert st.path == capi.PATH_RADIX_JOINCOUNT
        assert st.exchange_bytes == 0  # everything "sent" to itself
        res = be.select(capi.make_plan([ta, tb], **JOIN))
        single = res.rows()
        res.free()
    oa, ob = oracle.OracleTable([I]), oracle.OracleTable([I])
    oa.append_columns([a], [an])
    ob.append_columns([b])
    _, cells, nulls = oracle.select(capi.make_plan([oa, ob], **JOIN))
    want = oracle.rows_of(cells, nulls)
    assert helpers.canon(got) == helpers.canon(want)
    assert helpers.canon(single) == helpers.canon(want)


def uniform_specs(n):
    return ([capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=11, null_permille=20)],
            [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=100, hi=n + 4000, seed=12)])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_loopback_radix_exchange(world):
    """W virtual ranks on device 0: pass 1 per shard, k_radix_ship into the peers' arena slots, k_arena_barrier, pass 2 over
    W sources per partition.  Run twice so that both halves of the arena are used."""
    n = 1 << 22
    sa, sb = uniform_specs(n)
    for _ in range(2):
        parts = sharded_join([0] * world, sa, sb, n)
        assert all(p[2] == capi.PATH_RADIX_JOINCOUNT for p in parts)
        assert all(p[3] > 0 for p in parts)  # every rank pushed something to its peers
        assert check_against_oracle(parts) > 1000


@pytest.mark.parametrize("keys", ["zipf", "sorted", "heavy"])
def test_loopback_skewed_and_sorted_shards(keys):
    """what pass 1 cannot take - Zipf(1.1) keys, the reference README's auto-increment ids, one key with 5000 rows - in a
    DISTRIBUTED plan: every rank learns of the failure through the barrier flags and all of them answer with the
    direct-count path (owner reduction of the counters), exactly"""
    n = 1 << 20
    if keys == "zipf":
        sa = [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=(1 << 16) - 1, param=1.1, seed=21)]
        sb = [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=(1 << 16) - 1, param=1.1, seed=22)]
    elif keys == "sorted":
        sa = [capi.GenSpec(kind=capi.GEN_SEQUENCE, lo=1)]
        sb = [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=23)]
    else:
        sa = [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=255, seed=24)]  # 4096 rows per key: 8-bit counters wrap
        sb = [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=25)]
    parts = sharded_join([0, 0], sa, sb, n)
    assert all(p[2] == capi.PATH_DIRECT_COUNT for p in parts), [p[2] for p in parts]
    check_against_oracle(parts, histogram=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_in_process_multi_gpu_exchange(world):
    """one context per GPU in one process: the same exchange over real peer memory (NVLink)"""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    n = 1 << 24
    sa, sb = uniform_specs(n)
    parts = sharded_join(list(range(world)), sa, sb, n)
    assert all(p[2] == capi.PATH_RADIX_JOINCOUNT for p in parts)
    check_against_oracle(parts)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_torchrun_nccl_ipc_exchange(world):
    """one PROCESS per GPU (torchrun): NCCL rendezvous, arenas mapped with CUDA IPC - tests/dist_check.py"""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py"), "--log2-rows", "22"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0 and "DIST_CHECK OK world=%d" % world in out, out[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_loopback_scan_aggregate(world):
    """config 2 shape as a distributed plan: every rank scans its shard, the partials are all-gathered and folded in rank
    order, rank 0 returns the row.  COUNT / integer SUM / MIN / MAX exact, DOUBLE SUM / AVG 1e-9; two runs give the same bits."""
    n = 3 * (1 << 18)
    n_local = n // world
    lo, hi = 1 << 29, 3 * (1 << 29) - 1
    kw = dict(pred=[("col", 0, 0), ("int", lo), ("cmp", 6), ("col", 0, 0), ("int", hi), ("cmp", 5), ("and",)],
              out=[(OUT_COUNT_STAR,), (OUT_SUM, 0, 1), (capi.OUT_AVG, 0, 1), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1), (OUT_SUM, 0, 0)])
    specs = [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=(1 << 31) - 1, seed=41),
             capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, seed=42, null_permille=30)]

    def body(rank, be):
        t = be.create_table("T", [I, D])
        t.generate(n_local, specs, row_offset=rank * n_local)
        t.sync_stats()
        rows = []
        for _ in range(2):
            res = be.select(capi.make_plan([t], flags=PLAN_DISTRIBUTED, **kw))
            rows.append(res.rows())
            res.free()
        path = be.stats().path
        k, kv = t.read_column(0)
        v, vv = t.read_column(1)
        t.drop()
        return rows, path, k, v, vv

    parts = run_ranks([0] * world, body)
    assert all(p[1] == capi.PATH_SCAN_AGG for p in parts)
    assert all(p[0][0] == p[0][1] for p in parts)  # deterministic
    assert all(p[0][0] == [] for p in parts[1:]) and len(parts[0][0][0]) == 1  # rank 0 has the row
    k = np.concatenate([p[2] for p in parts])
    v = np.concatenate([p[3] for p in parts])
    vv = np.concatenate([p[4] for p in parts])
    ot = oracle.OracleTable([I, D])
    ot.append_columns([k, v], [None, (vv == 0).astype(np.uint8)])
    _, cells, nulls = oracle.select(capi.make_plan([ot], **kw))
    want = oracle.rows_of(cells, nulls)
    assert helpers.rows_close(parts[0][0][0], [helpers.norm_row(r) for r in want], rel=1e-9)
    assert parts[0][0][0][0][0] == want[0][0] and parts[0][0][0][0][5] == want[0][5]  # COUNT and the integer SUM are exact


@pytest.mark.parametrize("world", [2, 4])
def test_loopback_star_join(world):
    """config 5 shape as a distributed plan: BOTH tables sharded; the ranks' direct tables are merged into the whole dimension,
    every rank probes with its fact shard, rank 0 folds the accumulators and returns the groups"""
    nf, nd = 1 << 21, 1 << 14
    dspecs = [capi.GenSpec(kind=capi.GEN_PERMUTATION, lo=0, hi=nd - 1, seed=51), capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=499, seed=52)]
    fspecs = [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=nd + 99, seed=53),  # some foreign keys have no partner
              capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=-(1 << 40), hi=1 << 40, seed=54, null_permille=20),
              capi.GenSpec(kind=capi.GEN_UNIFORM_DBL, seed=55)]
    kw = dict(joins=[((0, 0), (1, 0))], group=[(0, 1)],
              out=[(OUT_COLUMN, 0, 1), (OUT_MIN, 1, 1), (OUT_MAX, 1, 1), (OUT_COUNT_STAR,), (OUT_SUM, 1, 2)])

    def body(rank, be):
        td, tf = be.create_table("D", [I, I]), be.create_table("F", [I, I, D])
        td.generate(nd // world, dspecs, row_offset=rank * (nd // world))
        tf.generate(nf // world, fspecs, row_offset=rank * (nf // world))
        td.sync_stats()
        tf.sync_stats()
        res = be.select(capi.make_plan([td, tf], flags=PLAN_DISTRIBUTED, **kw))
        rows, path, sent = res.rows(), be.stats().path, be.stats().exchange_bytes
        res.free()
        cols = [td.read_column(0), td.read_column(1), tf.read_column(0), tf.read_column(1), tf.read_column(2)]
        td.drop()
        tf.drop()
        return rows, path, sent, cols

    parts = run_ranks([0] * world, body)
    assert all(p[1] == capi.PATH_DIRECT_STAR for p in parts) and all(p[2] > 0 for p in parts)
    assert all(p[0] == [] for p in parts[1:]) and len(parts[0][0]) > 100
    cat = lambda i, j: np.concatenate([p[3][i][j] for p in parts])  # noqa: E731
    od, of = oracle.OracleTable([I, I]), oracle.OracleTable([I, I, D])
    od.append_columns([cat(0, 0), cat(1, 0)])
    of.append_columns([cat(2, 0), cat(3, 0), cat(4, 0)], [None, (cat(3, 1) == 0).astype(np.uint8), None])
    _, cells, nulls = oracle.select(capi.make_plan([od, of], **kw))
    assert helpers.canon_close(parts[0][0], oracle.rows_of(cells, nulls), rel=1e-9)


@pytest.mark.parametrize("shape", ["int_key_all_aggs", "null_keys_where", "double_key", "two_keys", "no_group_general_pred", "nothing_qualifies",
                                   "having_order_limit"])
@pytest.mark.parametrize("world", [2, 3])
def test_loopback_group_by_partials(world, shape):
    """GROUP BY / aggregates over ONE sharded table as a distributed plan (mdb_dist_group.cu): every rank aggregates its shard,
    the partial groups are all-gathered, rank 0 merges them by key (AVG = merged SUM / merged COUNT) and returns the groups.
    Keys and integer aggregates exact, DOUBLE SUM / AVG 1e-9 against the oracle on the unsharded table."""
    rng = np.random.default_rng(61)
    n_local = 3 * (1 << 15) + 17
    n = n_local * world
    types = [I, D, I, I]
    cols = [rng.integers(-300, 700, n), rng.random(n), rng.integers(-(1 << 40), 1 << 40, n), rng.integers(0, 7, n)]
    nulls = [np.zeros(n, np.uint8), (rng.random(n) < 0.03).astype(np.uint8), (rng.random(n) < 0.01).astype(np.uint8), np.zeros(n, np.uint8)]
    aggs = [(OUT_COUNT_STAR,), (capi.OUT_COUNT_COL, 0, 1), (OUT_SUM, 0, 1), (capi.OUT_AVG, 0, 1), (OUT_MIN, 0, 2), (OUT_MAX, 0, 2),
            (OUT_SUM, 0, 2), (capi.OUT_AVG, 0, 2), (OUT_MIN, 0, 1), (OUT_MAX, 0, 1)]
    if shape == "int_key_all_aggs":
        kw = dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0)] + aggs)
    elif shape == "null_keys_where":
        nulls[0] = (rng.random(n) < 0.04).astype(np.uint8)
        kw = dict(group=[(0, 0)], out=[(OUT_COUNT_STAR,), (OUT_COLUMN, 0, 0), (capi.OUT_AVG, 0, 1), (OUT_SUM, 0, 2)],
                  pred=[("col", 0, 1), ("dbl", 0.5), ("cmp", 1), ("col", 0, 3), ("int", 2), ("int", 5), ("in", 2), ("or",)])
    elif shape == "double_key":
        types = [I, D, I, D]
        cols[3] = rng.integers(-20, 20, n) / 4.0  # 40 distinct DOUBLE keys, -0.0 never produced
        kw = dict(group=[(0, 3)], out=[(OUT_COLUMN, 0, 3), (OUT_SUM, 0, 1), (OUT_COUNT_STAR,), (OUT_MIN, 0, 0), (capi.OUT_AVG, 0, 2)])
    elif shape == "two_keys":
        kw = dict(group=[(0, 0), (0, 3)], out=[(OUT_COLUMN, 0, 3), (OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (OUT_MAX, 0, 2), (capi.OUT_AVG, 0, 1)])
    elif shape == "having_order_limit":
        # tail operators run on rank 0's complete result: HAVING COUNT(*) > 100 ORDER BY COUNT(*) DESC, key LIMIT 25 OFFSET 3
        kw = dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,), (capi.OUT_AVG, 0, 1)],
                  having=[("out", 1), ("int", 100), ("cmp", 2)], order=[(1, True), (0, False)], limit=25, offset=3)
    elif shape == "no_group_general_pred":
        # an OR predicate: not the fused scan's shape, so the aggregate without GROUP BY comes here
        kw = dict(out=aggs, pred=[("col", 0, 0), ("int", 0), ("cmp", 1), ("col", 0, 3), ("int", 3), ("cmp", 3), ("or",)])
    else:
        kw = dict(group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)], pred=[("col", 0, 0), ("int", 5000), ("cmp", 2)])

    def body(rank, be):
        lo, hi = rank * n_local, (rank + 1) * n_local
        t = be.create_table("T", types)
        t.append_columns([c[lo:hi] for c in cols], [x[lo:hi] for x in nulls])
        t.sync_stats()
        res = be.select(capi.make_plan([t], flags=PLAN_DISTRIBUTED, **kw))
        rows, st = res.rows(), be.stats()
        res.free()
        t.drop()
        return rows, st.path, st.exchange_bytes

    parts = run_ranks([0] * world, body)
    assert all(p[1] == capi.PATH_GENERAL for p in parts)
    assert all(p[0] == [] for p in parts[1:])  # rank 0 returns the groups
    ot = oracle.OracleTable(types)
    ot.append_columns(cols, nulls)
    _, cells, onulls = oracle.select(capi.make_plan([ot], **kw))
    want = oracle.rows_of(cells, onulls)
    if shape == "nothing_qualifies":
        assert want == [] and parts[0][0] == []
        return
    assert len(want) >= 1 and all(p[2] > 0 for p in parts)
    if shape == "having_order_limit":  # ordered by (COUNT(*) DESC, key): an exact order, compare row by row
        assert len(want) == 25
        assert helpers.rows_close([helpers.norm_row(r) for r in parts[0][0]], [helpers.norm_row(r) for r in want], rel=1e-9)
        return
    assert helpers.canon_close(parts[0][0], want, rel=1e-9)
    # keys, counts and integer aggregates are exact
    int_cols = [i for i, o in enumerate(kw["out"]) if o[0] in (OUT_COUNT_STAR, capi.OUT_COUNT_COL) or
                (o[0] in (OUT_COLUMN, OUT_SUM, OUT_MIN, OUT_MAX) and types[o[2]] == I)]
    proj = lambda rows: helpers.canon([tuple(r[i] for i in int_cols) for r in rows])  # noqa: E731
    assert proj(parts[0][0]) == proj([helpers.norm_row(r) for r in want])
