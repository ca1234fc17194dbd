"""hash_product / hash_any / hash_all: the oracle against the reference binary (CPU)
and the C-ABI aggregators against the oracle (GPU).  Reference: GroupedProductImpl
(kernels/hash_aggregate_numeric.cc:311-335), GroupedAnyImpl / GroupedAllImpl (kernels/hash_aggregate.cc:1232-1397)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import arrow_oracle as ora

from .util import SEED

OPTS = [dict(skip_nulls=True, min_count=1), dict(skip_nulls=False, min_count=1), dict(skip_nulls=True, min_count=100),
        dict(skip_nulls=False, min_count=0)]


def _data(n=5000, offset=0):
    rng = np.random.default_rng(SEED)
    m = n + offset
    keys = pa.array(rng.integers(0, 40, m), mask=rng.random(m) < 0.05).slice(offset)
    b = pa.array(rng.random(m) < 0.7, mask=rng.random(m) < 0.2).slice(offset)
    iv = pa.array(rng.integers(-3, 4, m, dtype=np.int64), mask=rng.random(m) < 0.2).slice(offset)
    u8 = pa.array(rng.integers(0, 4, m, dtype=np.uint8), mask=rng.random(m) < 0.2).slice(offset)
    fv = pa.array(rng.uniform(0.5, 1.5, m), mask=rng.random(m) < 0.2).slice(offset)
    return keys, b, iv, u8, fv


@pytest.mark.parametrize("opts", OPTS)
def test_oracle_any_all_product_vs_reference_binary(opts):
    import pyarrow.acero  # noqa: F401
    keys, b, iv, u8, fv = _data()
    o = pc.ScalarAggregateOptions(**opts)
    t = pa.table({"k": keys, "b": b, "i": iv, "u": u8, "f": fv}).group_by("k", use_threads=False).aggregate(
        [("b", "any", o), ("b", "all", o), ("i", "product", o), ("u", "product", o), ("f", "product", o)]).sort_by("k")
    uniq, outs = ora.group_by([keys], [("hash_any", b, opts), ("hash_all", b, opts), ("hash_product", iv, opts),
                                       ("hash_product", u8, opts), ("hash_product", fv, opts)])
    mine = pa.table({"k": uniq[0], "b_any": outs[0], "b_all": outs[1], "i_product": outs[2], "u_product": outs[3],
                     "f_product": outs[4]}).sort_by("k")
    for c in ("k", "b_any", "b_all", "i_product", "u_product"):
        assert mine[c].combine_chunks().equals(t[c].combine_chunks()), (c, opts)
    a, w = mine["f_product"].combine_chunks(), t["f_product"].combine_chunks()
    assert a.is_valid().equals(w.is_valid())
    np.testing.assert_allclose(a.fill_null(0).to_numpy(), w.fill_null(0).to_numpy(), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("opts", OPTS)
@pytest.mark.parametrize("offset", [0, 5])
def test_gpu_any_all_product_vs_oracle(ctx, opts, offset):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    keys, b, iv, u8, fv = _data(60000, offset)
    dk = DeviceArray.from_arrow(keys, ctx)
    aggs = [("hash_any", b), ("hash_all", b), ("hash_product", iv), ("hash_product", u8), ("hash_product", fv)]
    uniq, outs = bc.group_by([dk], [(fn, DeviceArray.from_arrow(v, ctx), opts) for fn, v in aggs])
    ouniq, oouts = ora.group_by([keys], [(fn, v, opts) for fn, v in aggs])
    assert uniq[0].to_arrow().equals(ouniq[0])
    for (fn, v), got, want in zip(aggs, outs, oouts):
        got = got.to_arrow()
        if pa.types.is_floating(v.type):
            assert got.is_valid().equals(want.is_valid())
            np.testing.assert_allclose(got.fill_null(0).to_numpy(), want.fill_null(0).to_numpy(), rtol=1e-9)
        else:
            assert got.equals(want), (fn, str(v.type), opts)


@pytest.mark.gpu
def test_gpu_product_any_all_merge(ctx):
    """HashAggregateKernel::merge (kernel.h:720-725) for the new kinds: two halves consumed separately, merged through a
    group-id mapping, must equal one aggregator that saw every row"""
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    keys, b, iv, _, _ = _data(40000)
    n = len(keys)
    for fn, vals in (("hash_product", iv), ("hash_any", b), ("hash_all", b)):
        whole_u, (whole,) = bc.group_by([DeviceArray.from_arrow(keys, ctx)], [(fn, DeviceArray.from_arrow(vals, ctx), None)])
        g1, g2 = bc.Grouper([keys.type], ctx), bc.Grouper([keys.type], ctx)
        a1, a2 = bc.HashAggregator(fn, vals.type, ctx=ctx), bc.HashAggregator(fn, vals.type, ctx=ctx)
        for g, a, lo, hi in ((g1, a1, 0, n // 2), (g2, a2, n // 2, n)):
            ids = g.consume(DeviceArray.from_arrow(keys.slice(lo, hi - lo), ctx))
            a.resize(g.num_groups)
            a.consume(DeviceArray.from_arrow(vals.slice(lo, hi - lo), ctx), ids)
        mapping = g1.consume(g2.get_uniques()[0])     # GroupByNode::Merge: re-consume the other side's uniques
        a1.resize(g1.num_groups)
        a1.merge(a2, mapping)
        merged = pa.table({"k": g1.get_uniques()[0].to_arrow(), "v": a1.finalize().to_arrow()}).sort_by("k")
        want = pa.table({"k": whole_u[0].to_arrow(), "v": whole.to_arrow()}).sort_by("k")
        assert merged.equals(want), fn


# ---- hash_count_distinct (GroupedCountDistinctImpl, kernels/hash_aggregate.cc:1400-1478) ----------------------------
def _distinct_data(n, offset=0):
    rng = np.random.default_rng(SEED + 9)
    m = n + offset
    keys = pa.array(rng.integers(0, 30, m), mask=rng.random(m) < 0.05).slice(offset)
    i64 = pa.array(rng.integers(-5, 6, m, dtype=np.int64) * (1 << 40), mask=rng.random(m) < 0.1).slice(offset)
    u8 = pa.array(rng.integers(0, 7, m, dtype=np.uint8), mask=rng.random(m) < 0.1).slice(offset)
    f64 = pa.array(rng.integers(0, 9, m).astype(np.float64) / 4, mask=rng.random(m) < 0.1).slice(offset)
    words = [None if rng.random() < 0.1 else "w" * int(k % 4) + str(int(k)) for k in rng.integers(0, 12, m)]
    st = pa.array(words, pa.string()).slice(offset)
    return keys, [i64, u8, f64, st]


@pytest.mark.parametrize("mode", ["only_valid", "only_null", "all"])
def test_oracle_count_distinct_vs_reference_binary(mode):
    keys, cols = _distinct_data(4000, 3)
    o = pc.CountOptions(mode=mode)
    names = [f"c{j}" for j in range(len(cols))]
    t = pa.table([keys] + cols, names=["k"] + names).group_by("k", use_threads=False).aggregate(
        [(nm, "count_distinct", o) for nm in names]).sort_by("k")
    uniq, outs = ora.group_by([keys], [("hash_count_distinct", c, dict(mode=mode)) for c in cols])
    mine = pa.table([uniq[0]] + outs, names=["k"] + [nm + "_count_distinct" for nm in names]).sort_by("k")
    for c in mine.column_names:
        assert mine[c].combine_chunks().equals(t[c].combine_chunks()), (c, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["only_valid", "only_null", "all"])
@pytest.mark.parametrize("offset", [0, 5])
def test_gpu_count_distinct_vs_oracle(ctx, mode, offset):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    keys, cols = _distinct_data(50000, offset)
    dk = DeviceArray.from_arrow(keys, ctx)
    uniq, outs = bc.group_by([dk], [("hash_count_distinct", DeviceArray.from_arrow(c, ctx), dict(mode=mode)) for c in cols])
    ouniq, oouts = ora.group_by([keys], [("hash_count_distinct", c, dict(mode=mode)) for c in cols])
    assert uniq[0].to_arrow().equals(ouniq[0])
    for c, got, want in zip(cols, outs, oouts):
        assert got.to_arrow().equals(want), (c.type, mode)
        assert got.null_count == 0


@pytest.mark.gpu
def test_gpu_count_distinct_merge(ctx):
    """two partial states over different group numberings merged through a group_id_mapping = one state over the
    concatenated input (HashAggregateKernel::merge, kernel.h:757; GroupedCountDistinctImpl::Merge :1418-1439)"""
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    keys, cols = _distinct_data(30000)
    half = len(keys) // 2
    for col in cols:
        g_all = bc.Grouper([keys.type], ctx)
        whole = bc.HashAggregator("hash_count_distinct", col.type, mode="all", ctx=ctx)
        ids = g_all.consume(DeviceArray.from_arrow(keys, ctx))
        whole.resize(g_all.num_groups)
        whole.consume(DeviceArray.from_arrow(col, ctx), ids)
        want = whole.finalize().to_arrow()
        # partial A over the first half (its own grouper), partial B over the second half
        ga, gb = bc.Grouper([keys.type], ctx), bc.Grouper([keys.type], ctx)
        a = bc.HashAggregator("hash_count_distinct", col.type, mode="all", ctx=ctx)
        b = bc.HashAggregator("hash_count_distinct", col.type, mode="all", ctx=ctx)
        ia = ga.consume(DeviceArray.from_arrow(keys.slice(0, half), ctx))
        a.resize(ga.num_groups)
        a.consume(DeviceArray.from_arrow(col.slice(0, half), ctx), ia)
        ib = gb.consume(DeviceArray.from_arrow(keys.slice(half), ctx))
        b.resize(gb.num_groups)
        b.consume(DeviceArray.from_arrow(col.slice(half), ctx), ib)
        mapping = ga.consume(gb.get_uniques()[0])        # B's groups in A's numbering (new ones appended)
        a.resize(ga.num_groups)
        a.merge(b, mapping)
        got = a.finalize().to_arrow()
        # A's numbering is a permutation of the whole-input numbering: compare through the keys
        ka, kw = ga.get_uniques()[0].to_arrow(), g_all.get_uniques()[0].to_arrow()
        ta = pa.table({"k": ka, "c": got}).sort_by("k")
        tw = pa.table({"k": kw, "c": want}).sort_by("k")
        assert ta.equals(tw), col.type
