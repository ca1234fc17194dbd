"""The partitioned (hash radix-partition + shared-memory pre-aggregation) path of the fused
group-by, which only engages from 2^21 rows: 0, 1 and 2 partition passes, null keys, null values,
the all-ones key, several consume() batches, and growth past the cardinality hint -- against the
reference engine (Acero via pyarrow, rows sorted by key as its own tests do)."""
import numpy as np
import pyarrow as pa
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray

pytestmark = pytest.mark.gpu
SEED = 0x0FF1CE


def reference(keys, vals):
    import pyarrow.acero  # noqa: F401
    t = pa.table({"k": keys, "v": vals}).group_by("k", use_threads=False).aggregate([("v", "sum"), ("v", "count")])
    return t.sort_by("k")


def check(ctx, keys, vals, hint, batches=1):
    g = bc.GroupBySumCount(keys.type, vals.type, expected_groups=hint, ctx=ctx)
    n = len(keys)
    dk, dv = DeviceArray.from_arrow(keys, ctx), DeviceArray.from_arrow(vals, ctx)
    for b in range(batches):
        lo, hi = b * n // batches, (b + 1) * n // batches
        g.consume(dk.slice(lo, hi - lo), dv.slice(lo, hi - lo))
    k, s, c = [x.to_arrow() for x in g.finalize()]
    got = pa.table({"k": k, "v_sum": s, "v_count": c}).sort_by("k")
    want = reference(keys, vals)
    assert got["k"].combine_chunks().equals(want["k"].combine_chunks())
    assert got["v_count"].combine_chunks().equals(want["v_count"].combine_chunks())
    gs, ws = got["v_sum"].combine_chunks(), want["v_sum"].combine_chunks()
    if pa.types.is_integer(vals.type):
        assert gs.equals(ws)
    else:
        assert gs.is_valid().equals(ws.is_valid())
        np.testing.assert_allclose(gs.fill_null(0).to_numpy(), ws.fill_null(0).to_numpy(), rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("groups,hint", [(100, 100), (50_000, 50_000), (600_000, 600_000), (600_000, 0), (600_000, 1000)])
def test_partitioned_group_by_int64(ctx, groups, hint):
    n = 3_000_000
    rng = np.random.default_rng(SEED + groups)
    k = rng.integers(0, groups, n, dtype=np.int64)
    k[:3] = [-1, np.iinfo(np.int64).min, np.iinfo(np.int64).max]   # 0xFFFF.. is the table's empty marker
    keys = pa.array(k, mask=rng.random(n) < 0.01)
    vals = pa.array(rng.integers(-100, 100, n, dtype=np.int64), mask=rng.random(n) < 0.1)
    check(ctx, keys, vals, hint)


def test_partitioned_group_by_types_and_batches(ctx):
    n = 2_500_000
    rng = np.random.default_rng(SEED)
    keys32 = pa.array(rng.integers(-70000, 70000, n, dtype=np.int32), mask=rng.random(n) < 0.02)
    fvals = pa.array(rng.uniform(-1, 1, n), pa.float64(), mask=rng.random(n) < 0.1)
    check(ctx, keys32, fvals, 140_000)
    u16 = pa.array(rng.integers(0, 60000, n, dtype=np.uint16))
    f32 = pa.array(rng.uniform(0, 10, n).astype(np.float32), mask=rng.random(n) < 0.5)
    check(ctx, u16, f32, 0)
    # a group whose every value is null must exist with sum = null, count = 0
    k = rng.integers(0, 500_000, n, dtype=np.int64)
    v = rng.integers(-5, 5, n, dtype=np.int64)
    mask = (k % 7 == 0) | (rng.random(n) < 0.05)
    check(ctx, pa.array(k), pa.array(v, mask=mask), 500_000)
    # several consume() calls, each above the partitioning threshold; growing past the hint
    k2 = pa.array(rng.integers(0, 900_000, 3 * n, dtype=np.int64))
    v2 = pa.array(rng.integers(-100, 100, 3 * n, dtype=np.int64), mask=rng.random(3 * n) < 0.1)
    check(ctx, k2, v2, 10_000, batches=3)
