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
