#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check.py [--log2-rows 22]

Every rank loads its shard of the same global A and B (counter-based generator, so the global tables do not depend
on N), runs the README query as a MDBCU_PLAN_DISTRIBUTED plan and rank 0 checks the concatenation of all ranks'
groups against the CPU oracle run on the whole tables.  Prints one line 'DIST_CHECK OK ...' or raises.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import Dist, dist_env  # noqa: E402
from midoridb_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-rows", type=int, default=22)
    args = ap.parse_args()
    rank, world, local = dist_env()
    dist = Dist(rank, world)
    be = capi.Backend(local)
    uid = be.comm_unique_id() if rank == 0 else bytes(128)
    be.comm_init(rank, world, dist.bcast_bytes(uid, 128))

    n = 1 << args.log2_rows
    n_local = n // world
    I = capi.CT_INTEGER
    ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
    ta.generate(n_local, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=11, null_permille=20)], row_offset=rank * n_local)
    tb.generate(n_local, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=100, hi=n + 4000, seed=12)], row_offset=rank * n_local)
    ta.sync_stats()
    tb.sync_stats()
    kw = dict(joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)])
    res = be.select(capi.make_plan([ta, tb], flags=capi.PLAN_DISTRIBUTED, **kw))
    st = be.stats()
    (keys, cnts), _ = res.fetch_columns()
    res.free()
    a, av = ta.read_column(0)
    b, bv = tb.read_column(0)

    # gather shards and results on rank 0 (gloo, CPU tensors)
    import torch
    import torch.distributed as td

    def gather(arr):
        arr = np.ascontiguousarray(arr)
        if world == 1:
            return [arr]
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        td.all_gather(sizes, torch.tensor([arr.size], dtype=torch.int64))
        m = int(max(int(s[0]) for s in sizes))
        pad = np.zeros(m, dtype=arr.dtype)
        pad[:arr.size] = arr
        bufs = [torch.zeros(m, dtype=torch.from_numpy(pad).dtype) for _ in range(world)]
        td.all_gather(bufs, torch.from_numpy(pad))
        return [bufs[r].numpy()[:int(sizes[r][0])] for r in range(world)]

    all_keys, all_cnts = gather(keys), gather(cnts)
    all_a, all_av, all_b, all_bv = gather(a), gather(av.astype(np.int64)), gather(b), gather(bv.astype(np.int64))
    sent = gather(np.array([st.exchange_bytes], dtype=np.int64))
    if rank == 0:
        from oracle import oracle
        ga, gav = np.concatenate(all_a), np.concatenate(all_av)
        gb, gbv = np.concatenate(all_b), np.concatenate(all_bv)
        oa, ob = oracle.OracleTable([I]), oracle.OracleTable([I])
        oa.append_columns([ga], [(gav == 0).astype(np.uint8)])
        ob.append_columns([gb], [(gbv == 0).astype(np.uint8)])
        _, cells, _ = oracle.select(capi.make_plan([oa, ob], **kw))
        want = sorted(zip(cells[0].tolist(), cells[1].tolist()))
        got = sorted(zip(np.concatenate(all_keys).tolist(), np.concatenate(all_cnts).tolist()))
        assert got == want, "distributed result differs from the oracle (%d vs %d groups)" % (len(got), len(want))
        # ranks own disjoint key ranges
        for r in range(world - 1):
            if all_keys[r].size and all_keys[r + 1].size:
                assert all_keys[r].max() < all_keys[r + 1].min()
        print("DIST_CHECK OK world=%d rows=2^%d groups=%d nvlink_bytes_sent_per_rank=%s" %
              (world, args.log2_rows, len(got), [int(s[0]) for s in sent]))
    ta.drop()
    tb.drop()
    be.close()
    dist.close()


if __name__ == "__main__":
    main()
