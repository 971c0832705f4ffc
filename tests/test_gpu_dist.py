"""GPU tests (-m gpu) of the multi-GPU code path on ONE device: a world of size 1 takes the distributed plan's
decisions (global statistics mandatory, ownership of all partitions, no peer to push to), so its result must equal
the single-GPU path and the oracle.
Real 2/4/8-GPU runs use tests/dist_check.py under torchrun (gpurun --gpus N)."""
import numpy as np
import pytest

from midoridb_b200 import capi
from midoridb_b200.capi import CT_INTEGER, OUT_COLUMN, OUT_COUNT_STAR, PLAN_DISTRIBUTED
from oracle import oracle
from tests import helpers

pytestmark = pytest.mark.gpu
I = CT_INTEGER


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
        kw = dict(joins=[((0, 0), (1, 0))], group=[(0, 0)], out=[(OUT_COLUMN, 0, 0), (OUT_COUNT_STAR,)])
        with pytest.raises(capi.MdbError):  # global statistics are mandatory for distributed plans
            be.select(capi.make_plan([ta, tb], flags=PLAN_DISTRIBUTED, **kw))
        ta.sync_stats()
        tb.sync_stats()
        res = be.select(capi.make_plan([ta, tb], flags=PLAN_DISTRIBUTED, **kw))
        st = be.stats()
        got = res.rows()
        res.free()
        assert st.path == capi.PATH_RADIX_JOINCOUNT
        assert st.exchange_bytes == 0  # everything "sent" to itself
        res = be.select(capi.make_plan([ta, tb], **kw))
        single = res.rows()
        res.free()
    oa, ob = oracle.OracleTable([I]), oracle.OracleTable([I])
    oa.append_columns([a], [an])
    ob.append_columns([b])
    _, cells, nulls = oracle.select(capi.make_plan([oa, ob], **kw))
    want = oracle.rows_of(cells, nulls)
    assert helpers.canon(got) == helpers.canon(want)
    assert helpers.canon(single) == helpers.canon(want)
