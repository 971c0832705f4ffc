"""CPU test of the multi-process host logic (world_size 2, gloo): rendezvous, broadcast of the NCCL unique id,
max/sum over ranks as bench.py and tests/dist_check.py use them, and the partition-ownership arithmetic the
distributed radix join relies on (rank r owns partitions [r*P/W, (r+1)*P/W))."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    from bench import Dist, dist_env
    rank, world, local = dist_env()
    d = Dist(rank, world)
    uid = bytes(range(128)) if rank == 0 else bytes(128)
    got = d.bcast_bytes(uid, 128)
    assert got == bytes(range(128)), got
    assert d.max(10.0 + rank) == 10.0 + world - 1
    assert d.sum(rank + 1) == world * (world + 1) / 2
    d.barrier()
    # ownership: contiguous, disjoint, covering
    P = 3571
    lo, hi = rank * P // world, (rank + 1) * P // world
    total = d.sum(hi - lo)
    assert total == P
    d.close()
    print("worker", rank, "ok")
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_gloo_rendezvous():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "worker %d ok" % rank in out


LAYOUT_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as td
    from bench import Dist, dist_env
    from midoridb_b200 import capi
    rank, world, local = dist_env()
    d = Dist(rank, world)
    # the library's own ownership / slot arithmetic (mdbcu_dist_describe uses the functions the kernels use)
    for nparts, rows in ((4096, 1 << 28), (3571, 1000003), (7, 5000), (1, 10)):
        for W in (world, 3, 8):
            lay = capi.dist_describe(nparts, W, rows)
            first = list(lay.part_first)[:W + 1]
            assert first[0] == 0 and first[W] == nparts and all(a <= b for a, b in zip(first, first[1:]))
            # the push kernel's owner-of-partition formula agrees with the ranges (every partition, every rank)
            for p in list(range(min(nparts, 300))) + [nparts - 1, nparts // 2]:
                o = capi.dist_owner(p, nparts, W)
                assert first[o] <= p < first[o + 1], (p, o, first)
            assert capi.dist_owner(nparts, nparts, W) == -1
            # a side's region holds the streams of ALL partitions: 256-byte aligned sections in the order main, tail, cursor pairs
            assert lay.region_main_off == 0 and lay.region_tail_off >= nparts * lay.stream_cap * 2
            assert lay.region_cursor_off >= lay.region_tail_off + nparts * lay.tail_cap * 2
            assert lay.region_bytes >= lay.region_cursor_off + 4096 * 2 * 4
            assert all(x %% 256 == 0 for x in (lay.region_tail_off, lay.region_cursor_off, lay.region_bytes))
            assert lay.arena_half_bytes == 2 * lay.region_bytes
            assert lay.stream_cap %% 64 == 0 and lay.tail_cap %% 16 == 0
    # both ranks of THIS world derive the same layout and disjoint, covering ownership
    lay = capi.dist_describe(4096, world, 1 << 28)
    mine = torch.tensor([lay.part_first[rank], lay.part_first[rank + 1], lay.stream_cap, lay.tail_cap, lay.region_bytes], dtype=torch.int64)
    everyone = [torch.zeros(5, dtype=torch.int64) for _ in range(world)]
    td.all_gather(everyone, mine)
    assert all(int(e[2]) == lay.stream_cap and int(e[3]) == lay.tail_cap and int(e[4]) == lay.region_bytes for e in everyone)
    assert int(everyone[0][0]) == 0 and int(everyone[-1][1]) == 4096
    assert all(int(everyone[r][1]) == int(everyone[r + 1][0]) for r in range(world - 1))
    d.close()
    print("layout worker", rank, "ok")
""") % ROOT


def test_two_rank_partition_ownership_and_slot_layout():
    """the ownership / arena-slot arithmetic of the distributed radix join (mdb_radix.cu, mdb_radix_dist.cuh) through the
    library's host entry points, on two gloo ranks"""
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", LAYOUT_WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "layout worker %d ok" % rank in out
