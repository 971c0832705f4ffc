import sys, os, threading, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from midoridb_b200 import capi
I = capi.CT_INTEGER
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bes = [capi.Backend(d) for d in range(W)]
capi.comm_init_local(bes)
print("group ok", flush=True)
n = 1 << 22
def body(r):
    try:
        be = bes[r]
        ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
        ta.generate(n // W, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=1)], row_offset=r * (n // W))
        tb.generate(n // W, [capi.GenSpec(kind=capi.GEN_UNIFORM_INT, lo=0, hi=n - 1, seed=2)], row_offset=r * (n // W))
        print(r, "generated", flush=True)
        ta.sync_stats(); tb.sync_stats()
        print(r, "synced", flush=True)
        res = be.select(capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)], flags=capi.PLAN_DISTRIBUTED))
        print(r, "rows", res.nrows, "path", be.stats().path, flush=True)
    except BaseException:
        traceback.print_exc()
ts = [threading.Thread(target=body, args=(r,)) for r in range(W)]
[t.start() for t in ts]; [t.join() for t in ts]
