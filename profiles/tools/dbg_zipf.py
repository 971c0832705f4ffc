import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from midoridb_b200 import capi
I = capi.CT_INTEGER
be = capi.Backend(0)
n = 1 << 20
ta, tb = be.create_table("A", [I]), be.create_table("B", [I])
ta.generate(n, [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=(1 << 16) - 1, param=1.1, seed=21)])
tb.generate(n, [capi.GenSpec(kind=capi.GEN_ZIPF, lo=0, hi=(1 << 16) - 1, param=1.1, seed=22)])
print("generated", flush=True)
t0 = time.time()
res = be.select(capi.make_plan([ta, tb], joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(capi.OUT_COLUMN, 0, 0), (capi.OUT_COUNT_STAR,)]))
print("path", be.stats().path, "rows", res.nrows, "s", time.time() - t0, flush=True)
