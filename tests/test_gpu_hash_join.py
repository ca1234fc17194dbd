"""b2_hash_join (csrc/hash_join.cu): the matching row pairs of an equi-join, against the oracle (exact order: left rows in
row order, their matches in right-row order) and against the reference binary's HashJoinNode through pyarrow.Table.join
(order unspecified there: compared as sorted pairs).  A null key matches nothing (JoinKeyCmp::EQ, acero/options.h:384-392)."""
import numpy as np
import pyarrow as pa
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from oracle import arrow_oracle as ora
from tests.test_gpu_grouper_wide import words
from tests.util import SEED, random_array

pytestmark = pytest.mark.gpu
JOIN_TYPES = ["inner", "left outer", "left semi", "left anti", "full outer"]


def dev(arr, ctx):
    return DeviceArray.from_arrow(arr, ctx)


def reference_pairs(left_keys, right_keys, join_type):
    """(left row, right row or None) pairs from the reference binary: join tables that carry their row numbers"""
    names = [f"k{j}" for j in range(len(left_keys))]
    lt = pa.table(list(left_keys) + [pa.array(np.arange(len(left_keys[0]), dtype=np.int64))], names=names + ["lrow"])
    rt = pa.table(list(right_keys) + [pa.array(np.arange(len(right_keys[0]), dtype=np.int64))], names=names + ["rrow"])
    out = lt.join(rt, keys=names, join_type=join_type, use_threads=False)
    lrow = [(-1 if l is None else l) for l in out["lrow"].to_pylist()]
    rrow = out["rrow"].to_pylist() if "rrow" in out.column_names else [None] * len(lrow)
    return sorted(zip(lrow, [(-1 if r is None else r) for r in rrow]))


def check(ctx, left_keys, right_keys, join_type):
    got_l, got_r = bc.hash_join_indices([dev(k, ctx) for k in left_keys], [dev(k, ctx) for k in right_keys], join_type)
    want_l, want_r = ora.hash_join_indices(left_keys, right_keys, join_type)
    assert got_l.to_arrow().equals(want_l), join_type
    if want_r is None:
        assert got_r is None
        mine = sorted((l, -1) for l in got_l.to_arrow().to_pylist())
    else:
        assert got_r.to_arrow().equals(want_r), join_type
        assert got_r.null_count == want_r.null_count
        assert got_l.null_count == want_l.null_count
        mine = sorted(zip([(-1 if l is None else l) for l in got_l.to_arrow().to_pylist()],
                          [(-1 if r is None else r) for r in got_r.to_arrow().to_pylist()]))
    ref = reference_pairs(left_keys, right_keys, join_type)
    if join_type in ("left semi", "left anti"):
        ref = sorted((l, -1) for l, _ in ref)
    assert mine == ref, join_type


@pytest.mark.parametrize("join_type", JOIN_TYPES)
@pytest.mark.parametrize("kt", [pa.int64(), pa.int32(), pa.float64(), pa.string()], ids=str)
def test_single_key(ctx, kt, join_type):
    def col(n, seed, off):
        if pa.types.is_string(kt):
            return words(kt, n, 60, 0.1, seed, offset=off)
        return random_array(kt, n, 0.1, seed, lo=0, hi=80, offset=off)
    check(ctx, [col(9000, SEED + 1, 3)], [col(5000, SEED + 2, 0)], join_type)     # many-to-many with nulls on both sides
    check(ctx, [col(100, SEED + 3, 0)], [col(0, SEED, 0)], join_type)            # empty build side
    check(ctx, [col(0, SEED, 0)], [col(100, SEED + 4, 0)], join_type)            # empty probe side


@pytest.mark.parametrize("join_type", JOIN_TYPES)
def test_multi_column_and_wide_keys(ctx, join_type):
    rng = np.random.default_rng(SEED)

    def side(n, seed):
        r = np.random.default_rng(seed)
        return [pa.array(r.integers(0, 6, n) * (1 << 40), pa.int64(), mask=r.random(n) < 0.05),
                words(pa.string(), n, 5, 0.05, seed),
                pa.array(r.integers(0, 3, n).astype(np.int16), pa.int16())]
    check(ctx, side(6000, SEED + 10), side(4000, SEED + 11), join_type)
    # unique keys on the build side (the dimension-table shape), every probe row matches exactly once or not at all
    dim = [pa.array(rng.permutation(50000).astype(np.int64))]
    fact = [pa.array(rng.integers(0, 60000, 200000).astype(np.int64))]
    check(ctx, fact, dim, join_type)


def test_join_payload_through_take(ctx):
    """the caller's materialize step: payload columns gathered with take() through the returned indices == Table.join"""
    rng = np.random.default_rng(SEED + 5)
    n_l, n_r = 30000, 2000
    lk = pa.array(rng.integers(0, 2500, n_l), pa.int64())
    rk = pa.array(rng.permutation(2500)[:n_r].astype(np.int64))
    lv = pa.array(rng.normal(size=n_l))
    rv = pa.array(rng.integers(0, 1000, n_r), pa.int32())
    li, ri = bc.hash_join_indices([dev(lk, ctx)], [dev(rk, ctx)], "left outer")
    got = pa.table({"k": bc.take(dev(lk, ctx), li).to_arrow(), "lv": bc.take(dev(lv, ctx), li).to_arrow(),
                    "rv": bc.take(dev(rv, ctx), ri).to_arrow()})
    ref = pa.table({"k": lk, "lv": lv}).join(pa.table({"k": rk, "rv": rv}), keys="k", join_type="left outer", use_threads=False)
    order = [("k", "ascending"), ("lv", "ascending")]
    assert got.sort_by(order).equals(ref.select(["k", "lv", "rv"]).sort_by(order))
