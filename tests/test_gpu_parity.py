"""Parity of the CUDA path (through the C-ABI, via arrow_b200.compute) against the oracle,
the reference's known-answer vectors and -- since the same image ships it -- the
reference binary (pyarrow.compute) on identical inputs.  Bit-exact for integer, index,
selection, sort and cast outputs and for single IEEE float ops; float SUMS use rtol 1e-9
(summation order differs; the reference itself compares those approximately,
acero/hash_aggregate_test.cc:3641-3663)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from oracle import arrow_oracle as ora
from tests.util import INT_TYPES, NUMERIC_TYPES, SEED, TYPE_BY_NAME, assert_equal, equal_nan, from_json, kat, random_array

pytestmark = pytest.mark.gpu
KAT = kat()


def dev(arr, ctx):
    return DeviceArray.from_arrow(arr, ctx)


# ---------------------------------------------------------------- known answers
@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_kat_filter(ctx, t):
    for c in KAT["filter_numeric_basics"]["cases"]:
        v, m = from_json(t, c["values"]), from_json(pa.bool_(), c["filter"])
        for ns in ("emit_null", "drop"):
            got = bc.filter(dev(v, ctx), dev(m, ctx), ns).to_arrow()
            assert_equal(got, from_json(t, c[ns]), f"{ns} {c}")
    with pytest.raises(pa.ArrowInvalid, match="same length"):
        bc.filter(dev(from_json(t, [7, 8, 9]), ctx), dev(from_json(pa.bool_(), []), ctx))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_kat_take(ctx, t):
    for c in KAT["take_numeric_basics"]["cases"]:
        for it in (pa.int8(), pa.uint32(), pa.int64()):
            got = bc.take(dev(from_json(t, c["values"]), ctx), dev(from_json(it, c["indices"]), ctx)).to_arrow()
            assert_equal(got, from_json(t, c["expected"]), str(c))
    for c in KAT["take_numeric_basics"]["index_errors"]:
        with pytest.raises(pa.ArrowIndexError, match="out of bounds"):
            bc.take(dev(from_json(t, c["values"]), ctx), dev(from_json(pa.int8(), c["indices"]), ctx))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_kat_sort(ctx, t):
    cases = list(KAT["sort_integral"]["cases"]) if pa.types.is_integer(t) else list(KAT["sort_real"]["cases"])
    if t == pa.int64():
        cases += KAT["sort_integral"]["int64_cases"]
    for c in cases:
        got = bc.array_sort_indices(dev(from_json(t, c["values"]), ctx), c["order"], c["null_placement"]).to_arrow()
        assert_equal(got, pa.array(c["expected"], pa.uint64()), str(c))


def test_kat_cast(ctx):
    for c in KAT["cast_float_to_float"]["cases"]:
        got = bc.cast(dev(from_json(TYPE_BY_NAME[c["from"]], c["values"]), ctx), TYPE_BY_NAME[c["to"]]).to_arrow()
        assert_equal(got, from_json(TYPE_BY_NAME[c["to"]], c["expected"]))
    for c in KAT["cast_int_to_float_bounds"]["ok"]:
        src = from_json(TYPE_BY_NAME[c["from"]], c["values"])
        assert_equal(bc.cast(dev(src, ctx), TYPE_BY_NAME[c["to"]]).to_arrow(), pc.cast(src, TYPE_BY_NAME[c["to"]]))
    for c in KAT["cast_int_to_float_bounds"]["fails"]:
        src = from_json(TYPE_BY_NAME[c["from"]], c["values"])
        with pytest.raises(pa.ArrowInvalid) as want:
            pc.cast(src, TYPE_BY_NAME[c["to"]])
        with pytest.raises(pa.ArrowInvalid) as got:
            bc.cast(dev(src, ctx), TYPE_BY_NAME[c["to"]])
        assert str(got.value) == str(want.value)
    c = KAT["cast_overflow_in_null_slot"]
    v = pa.array(c["values"], TYPE_BY_NAME[c["from"]], mask=~np.array(c["validity"], dtype=bool))
    assert_equal(bc.cast(dev(v, ctx), TYPE_BY_NAME[c["to"]]).to_arrow(), from_json(TYPE_BY_NAME[c["to"]], c["expected"]))


def _sorted_by_key(keys, cols):
    order = pc.sort_indices(keys, null_placement="at_end")
    return [pc.take(c, order) for c in [keys] + list(cols)]


def test_kat_group_by(ctx):
    c = KAT["group_by_count_only"]
    arg = pa.array([r[0] for r in c["rows"]], pa.float64())
    key = pa.array([r[1] for r in c["rows"]], pa.int64())
    for mode in ("only_valid", "only_null", "all"):
        uniq, (cnt,) = bc.group_by([dev(key, ctx)], [("hash_count", dev(arg, ctx), {"mode": mode})])
        k, v = _sorted_by_key(uniq[0].to_arrow(), [cnt.to_arrow()])
        assert k.to_pylist() == [r[0] for r in c[mode]] and v.to_pylist() == [r[1] for r in c[mode]]
    c = KAT["group_by_sum_only"]
    arg = pa.array([r[0] for r in c["rows"]], pa.float64())
    key = pa.array([r[1] for r in c["rows"]], pa.int64())
    uniq, (s,) = bc.group_by([dev(key, ctx)], [("hash_sum", dev(arg, ctx), None)])
    k, v = _sorted_by_key(uniq[0].to_arrow(), [s.to_arrow()])
    assert k.to_pylist() == [r[0] for r in c["expected"]] and v.to_pylist() == [r[1] for r in c["expected"]]
    # fused path, same vectors
    f = bc.GroupBySumCount(pa.int64(), pa.float64(), ctx=ctx)
    f.consume(dev(key, ctx), dev(arg, ctx))
    fk, fs, fc = [x.to_arrow() for x in f.finalize()]
    k, v = _sorted_by_key(fk, [fs])
    assert k.to_pylist() == [r[0] for r in c["expected"]] and v.to_pylist() == [r[1] for r in c["expected"]]
    c = KAT["grouper_int64"]
    g = bc.Grouper([pa.int64()], ctx)
    assert g.consume(dev(pa.array(c["keys"], pa.int64()), ctx)).to_arrow().to_pylist() == c["ids"]
    assert g.num_groups == 4


# ---------------------------------------------------------------- random, vs oracle and reference
@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("null_p", [0.0, 0.01, 0.1, 0.999, 1.0])
def test_filter_random(ctx, t, null_p):
    # FilterRandomTest, vector_selection_test.cc:2241-2259 (+ larger sizes that span many tiles)
    for i, (n, true_p, mask_null, off) in enumerate([(1024, 0.5, 0.0, 0), (1024, 0.1, 0.05, 3), (1024, 0.999, 0.5, 2),
                                                     (1024, 0.0, 0.0, 1), (1024, 1.0, 0.0, 0), (70001, 0.5, 0.1, 5),
                                                     (70001, 0.01, 0.0, 0)]):
        v = random_array(t, n, null_p, SEED + i, offset=off)
        m = random_array(pa.bool_(), n, mask_null, SEED + 77 + i, hi=true_p, offset=(off * 5) % 7)
        dv, dm = dev(v, ctx), dev(m, ctx)
        for ns in ("drop", "emit_null"):
            got = bc.filter(dv, dm, ns)
            want = ora.filter(v, m, ns)
            assert_equal(got.to_arrow(), want, f"{t} {ns} case {i}")
            assert got.null_count == want.null_count
            assert bc.filter_output_size(dm, ns) == len(want)
            assert_equal(got.to_arrow(), pc.filter(v, m, null_selection_behavior=ns))
        # ValidateFilter: Filter(v, f) == Take(v, GetTakeIndices(f))  (vector_selection_test.cc:352-373)
        if n <= 1024:
            for ns in ("drop", "emit_null"):
                idx = bc.take_indices_from_filter(dm, ns)
                assert_equal(idx.to_arrow(), ora.take_indices_from_filter(m, ns))
                assert_equal(bc.take(dv, idx).to_arrow(), bc.filter(dv, dm, ns).to_arrow())


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("it", [pa.int8(), pa.uint8(), pa.int16(), pa.uint16(), pa.int32(), pa.uint32(), pa.int64(),
                                pa.uint64()], ids=str)
def test_take_random(ctx, t, it):
    # TakeRandomTest, vector_selection_test.cc:2286-2315
    for null_p in (0.0, 0.001, 0.05, 0.25, 0.95, 1.0):
        v = random_array(t, 1025, null_p, SEED, offset=1)
        hi = min(1024, np.iinfo(it.to_pandas_dtype()).max)
        for n_idx, off in ((257, 3), (9000, 0)):
            idx = random_array(it, n_idx, null_p, SEED + 5, lo=0, hi=hi, offset=off)
            got = bc.take(dev(v, ctx), dev(idx, ctx))
            want = ora.take(v, idx)
            assert_equal(got.to_arrow(), want, f"{t} {it} {null_p}")
            assert got.null_count == want.null_count
            assert_equal(got.to_arrow(), pc.take(v, idx))


@pytest.mark.parametrize("it", [pa.int32(), pa.uint32(), pa.int64(), pa.uint64()], ids=str)
@pytest.mark.parametrize("band_kb", ["1", "2", "3"])
def test_take_validity_bands(ctx, it, band_kb, monkeypatch):
    """Big validity bitmaps are probed in L2-sized bands (selection_take.cu take_bands): the gather kernel probes band 0,
    follow-up launches clear the bits of the other bands.  B2_TAKE_BAND_KB forces the banded path at test sizes
    (1 KB of bitmap = 8192 rows per band); results must equal the unbanded kernel, the oracle and the reference."""
    monkeypatch.setenv("B2_TAKE_BAND_KB", band_kb)
    for t in (pa.float64(), pa.int32(), pa.uint8()):
        for n_v, n, vnull, inull, off in ((40000, 100003, 0.1, 0.0, 0), (33000, 70000, 0.5, 0.1, 3), (16384, 5, 0.9, 0.0, 1),
                                          (45000, 262144, 0.02, 0.02, 0)):
            v = random_array(t, n_v, vnull, SEED + n, offset=off)
            idx = random_array(it, n, inull, SEED + 7, lo=0, hi=n_v - 1, offset=off)
            got = bc.take(dev(v, ctx), dev(idx, ctx))
            want = ora.take(v, idx)
            assert_equal(got.to_arrow(), want, f"{t} {it} band={band_kb} n={n}")
            assert got.null_count == want.null_count
            assert_equal(got.to_arrow(), pc.take(v, idx))
    # the fused pipeline shares the band plan
    values = random_array(pa.float64(), 40000, 0.2, SEED, lo=-1000, hi=1000)
    idx = random_array(it, 150001, 0.05, SEED + 1, lo=0, hi=39999, offset=1)
    other = random_array(pa.float32(), 150001, 0.1, SEED + 2, lo=-10, hi=10, offset=2)
    got = bc.take_cast_arith(dev(values, ctx), dev(idx, ctx), pa.float32(), "add", dev(other, ctx))
    ref = pc.add(pc.cast(pc.take(values, idx), pa.float32(), safe=False), other)
    assert got.to_arrow().equals(ref) and got.null_count == ref.null_count
    # an out-of-range index still raises, with the follow-up launches in the queue
    bad = pa.array([1, 2, 40000, 3] * 10, it)
    with pytest.raises(pa.ArrowIndexError, match="Index 40000 out of bounds"):
        bc.take(dev(values, ctx), dev(bad, ctx))


def test_take_errors(ctx):
    v = dev(pa.array(range(100), pa.int64()), ctx)
    for bad, it in ((100, pa.int32()), (-1, pa.int64()), (2**40, pa.int64()), (200, pa.uint8())):
        idx = pa.array([1, 2, bad, 3], it)
        with pytest.raises(pa.ArrowIndexError) as want:
            pc.take(pa.array(range(100), pa.int64()), idx)
        with pytest.raises(pa.ArrowIndexError) as got:
            bc.take(v, dev(idx, ctx))
        assert str(got.value) == str(want.value)
    # out-of-range index under a null is fine
    idx = pa.array([1, 1000, 3], pa.int32(), mask=np.array([False, True, False]))
    assert_equal(bc.take(v, dev(idx, ctx)).to_arrow(), pc.take(pa.array(range(100), pa.int64()), idx))


@pytest.mark.parametrize("src", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("dst", NUMERIC_TYPES, ids=str)
def test_cast_random(ctx, src, dst):
    if src == dst:
        return
    for n, off in ((500, 1), (20000, 0), (4099, 7)):
        v = random_array(src, n, 0.1, SEED, lo=0, hi=100, offset=off)
        assert_equal(bc.cast(dev(v, ctx), dst, safe=False).to_arrow(), pc.cast(v, dst, safe=False), f"{src}->{dst}")
        if pa.types.is_floating(src) and pa.types.is_integer(dst):
            v = pc.round(v)
        got = bc.cast(dev(v, ctx), dst).to_arrow()
        assert_equal(got, pc.cast(v, dst))
        assert_equal(got, ora.cast_array(v, dst))
    # full-range inputs: wrap when unsafe / identical error text when safe
    wide = random_array(src, 3000, 0.1, SEED + 1)
    if pa.types.is_integer(src) or pa.types.is_floating(dst):
        assert_equal(bc.cast(dev(wide, ctx), dst, safe=False).to_arrow(), pc.cast(wide, dst, safe=False))
    try:
        want = pc.cast(wide, dst)
    except pa.ArrowInvalid as e:
        with pytest.raises(pa.ArrowInvalid) as got:
            bc.cast(dev(wide, ctx), dst)
        assert str(got.value) == str(e), f"{src}->{dst}"
    else:
        assert_equal(bc.cast(dev(wide, ctx), dst).to_arrow(), want)


def test_cast_float_edge_values(ctx):
    # denormals, +-0, inf, nan, round-to-nearest-even ties: f64 -> f32 must be bit-exact
    vals = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 1e-40, -1e-46, 3.4028235e38, 3.4028236e38, 1e39,
                     1.0 + 2**-24, 1.0 + 2**-23 + 2**-24, 16777217.0, 0.1, 1 / 3], dtype=np.float64)
    v = pa.array(vals, pa.float64())
    got = bc.cast(dev(v, ctx), pa.float32(), safe=False).to_arrow()
    want = pc.cast(v, pa.float32(), safe=False)
    assert got.buffers()[1].to_pybytes()[:4 * len(vals)] == want.buffers()[1].to_pybytes()[:4 * len(vals)]
    # float -> int truncation errors: same first offending value, same text
    for dst in (pa.int32(), pa.uint8(), pa.int64()):
        x = pa.array([1.0, 2.0, 3.5, 4.5], pa.float64())
        with pytest.raises(pa.ArrowInvalid) as want_e:
            pc.cast(x, dst)
        with pytest.raises(pa.ArrowInvalid) as got_e:
            bc.cast(dev(x, ctx), dst)
        assert str(got_e.value) == str(want_e.value)


ARITH = ["add", "subtract", "multiply", "divide", "add_checked", "subtract_checked", "multiply_checked", "divide_checked"]


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("op", ARITH)
def test_arithmetic_random(ctx, t, op):
    small = dict(lo=1, hi=11) if pa.types.is_integer(t) else dict(lo=-100, hi=100)
    cases = [(random_array(t, 300, 0.1, SEED, offset=1, **small), random_array(t, 300, 0.1, SEED + 1, offset=2, **small)),
             (random_array(t, 40000, 0.1, SEED + 4, **small), random_array(t, 40000, 0.0, SEED + 5, **small))]
    if not op.endswith("_checked") and "divide" not in op:
        cases.append((random_array(t, 5000, 0.1, SEED + 2), random_array(t, 5000, 0.0, SEED + 3)))  # wraps
    for a, b in cases:
        da, db = dev(a, ctx), dev(b, ctx)
        try:
            want = getattr(pc, op)(a, b)
        except pa.ArrowInvalid as e:
            with pytest.raises(pa.ArrowInvalid) as got:
                getattr(bc, op)(da, db)
            assert str(got.value) == str(e)
            continue
        got = getattr(bc, op)(da, db)
        assert equal_nan(got.to_arrow(), want), f"{t} {op}"
        assert got.null_count == want.null_count
        assert equal_nan(getattr(bc, op)(da, b[0]).to_arrow(), getattr(pc, op)(a, b[0]))   # ArrayScalar
        assert equal_nan(getattr(bc, op)(a[0], db).to_arrow(), getattr(pc, op)(a[0], b))   # ScalarArray
        assert equal_nan(getattr(bc, op)(da, pa.scalar(None, t)).to_arrow(), getattr(pc, op)(a, pa.scalar(None, t)))


def test_arithmetic_float_bit_exact(ctx):
    rng = np.random.default_rng(SEED)
    for t, dt in ((pa.float32(), np.float32), (pa.float64(), np.float64)):
        a = rng.standard_normal(100000).astype(dt) * dt(1e-20)
        b = rng.standard_normal(100000).astype(dt) * dt(1e20)
        a[:5] = [np.inf, -np.inf, np.nan, 0.0, -0.0]
        b[:5] = [-np.inf, 1.0, 1.0, -0.0, 0.0]
        tiny = np.finfo(dt).tiny
        a[5:8] = [tiny, tiny / 2, -tiny / 4]   # denormal results must not be flushed
        b[5:8] = [-tiny / 2, tiny / 4, tiny / 8]
        for op in ("add", "subtract", "multiply", "divide"):
            got = getattr(bc, op)(dev(pa.array(a, t), ctx), dev(pa.array(b, t), ctx)).to_arrow()
            want = getattr(pc, op)(pa.array(a, t), pa.array(b, t))
            # bit-exact wherever the result is a number; where it is NaN both must be NaN (the NaN
            # *payload* of an invalid operation is implementation-defined: x86 gives the negative
            # "real indefinite" quiet NaN, sm_100 the positive canonical one)
            g = np.frombuffer(got.buffers()[1], dtype=dt)[:len(a)]
            w = np.frombuffer(want.buffers()[1], dtype=dt)[:len(a)]
            assert np.array_equal(np.isnan(g), np.isnan(w)), f"{t} {op}"
            ok = ~np.isnan(w)
            ut = np.uint32 if dt == np.float32 else np.uint64
            assert np.array_equal(g.view(ut)[ok], w.view(ut)[ok]), f"{t} {op}"


def test_arithmetic_errors(ctx):
    for t in INT_TYPES:
        info = np.iinfo(t.to_pandas_dtype())
        a = pa.array([1, info.max, 1], t)
        for op, b in (("add_checked", [1, 2, 1]), ("multiply_checked", [1, 2, 1]), ("divide", [1, 0, 1]),
                      ("divide_checked", [1, 0, 1])):
            with pytest.raises(pa.ArrowInvalid) as want:
                getattr(pc, op)(a, pa.array(b, t))
            with pytest.raises(pa.ArrowInvalid) as got:
                getattr(bc, op)(dev(a, ctx), dev(pa.array(b, t), ctx))
            assert str(got.value) == str(want.value), f"{t} {op}"
        z = pa.array([0, 1, 1], t, mask=np.array([True, False, False]))
        assert_equal(bc.divide(dev(a, ctx), dev(z, ctx)).to_arrow(), pc.divide(a, z))
        assert_equal(bc.add_checked(dev(pa.array([info.max, 1, 2], t, mask=np.array([True, False, False])), ctx),
                                    dev(pa.array([5, 1, 2], t), ctx)).to_arrow(),
                     pc.add_checked(pa.array([info.max, 1, 2], t, mask=np.array([True, False, False])), pa.array([5, 1, 2], t)))
    m = pa.array([-128, 5], pa.int8())
    assert_equal(bc.divide(dev(m, ctx), dev(pa.array([-1, 2], pa.int8()), ctx)).to_arrow(), pc.divide(m, pa.array([-1, 2], pa.int8())))
    with pytest.raises(pa.ArrowInvalid, match="overflow"):
        bc.divide_checked(dev(m, ctx), dev(pa.array([-1, 2], pa.int8()), ctx))
    with pytest.raises(pa.ArrowInvalid, match="divide by zero"):
        bc.divide_checked(dev(pa.array([1.0]), ctx), dev(pa.array([0.0]), ctx))
    with pytest.raises(pa.ArrowInvalid, match="same length"):
        bc.add(dev(pa.array([1, 2]), ctx), dev(pa.array([1]), ctx))


def test_mixed_type_dispatch(ctx):
    pairs = [(pa.int8(), pa.uint8()), (pa.int32(), pa.uint32()), (pa.uint64(), pa.int8()), (pa.int64(), pa.float32()),
             (pa.uint16(), pa.float64()), (pa.int16(), pa.int64())]
    for ta, tb in pairs:
        a, b = random_array(ta, 1000, 0.1, SEED, lo=0, hi=50), random_array(tb, 1000, 0.1, SEED + 9, lo=0, hi=50)
        assert_equal(bc.add(dev(a, ctx), dev(b, ctx)).to_arrow(), pc.add(a, b), f"{ta}+{tb}")
        assert_equal(bc.less(dev(a, ctx), dev(b, ctx)).to_arrow(), pc.less(a, b), f"{ta}<{tb}")
    a = random_array(pa.int32(), 100, 0.1, SEED)
    assert_equal(bc.add(dev(a, ctx), 1.5).to_arrow(), pc.add(a, 1.5))
    assert_equal(bc.call_function("multiply", [dev(a, ctx), 3]).to_arrow(), pc.multiply(a, 3))
    with pytest.raises(pa.ArrowKeyError):
        bc.call_function("no_such_function", [dev(a, ctx)])


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_compare_random(ctx, t):
    for n, o1, o2 in ((777, 3, 5), (100000, 0, 0), (4096, 1, 0)):
        a = random_array(t, n, 0.1, SEED, lo=0, hi=20, offset=o1)
        b = random_array(t, n, 0.1, SEED + 1, lo=0, hi=20, offset=o2)
        if pa.types.is_floating(t):
            x = a.to_numpy(zero_copy_only=False).copy()
            x[::13] = np.nan
            a = pa.array(x, t, mask=~ora.validity(a))
        for op in ("equal", "not_equal", "greater", "greater_equal", "less", "less_equal"):
            got = getattr(bc, op)(dev(a, ctx), dev(b, ctx))
            assert_equal(got.to_arrow(), getattr(pc, op)(a, b), f"{t} {op}")
            assert_equal(got.to_arrow(), ora.compare(op, a, b))
            assert_equal(getattr(bc, op)(dev(a, ctx), b[1]).to_arrow(), getattr(pc, op)(a, b[1]))
            assert_equal(getattr(bc, op)(a[2], dev(b, ctx)).to_arrow(), getattr(pc, op)(a[2], b))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("null_p", [0.0, 0.1, 0.5, 1.0])
def test_sort_random(ctx, t, null_p):
    # ValidateSorted-style random test (vector_sort_test.cc:972-1061) against both oracles
    for i, (n, rng) in enumerate([(100, dict()), (1000, dict(lo=0, hi=10)), (50000, dict()), (50000, dict(lo=0, hi=300))]):
        v = random_array(t, n, null_p, SEED + i, offset=i % 3, **rng)
        if pa.types.is_floating(t):
            a = v.to_numpy(zero_copy_only=False).copy()
            a[::17] = np.nan
            a[5::31] = -0.0
            a[7::31] = 0.0
            a[11::97] = np.inf
            a[13::97] = -np.inf
            v = pa.array(a, t, mask=~ora.validity(v))
        dv = dev(v, ctx)
        for order in ("ascending", "descending"):
            for np_ in ("at_end", "at_start"):
                got = bc.array_sort_indices(dv, order, np_).to_arrow()
                assert_equal(got, ora.sort_indices(v, order, np_), f"{t} {order} {np_} n={n}")
                assert_equal(got, pc.array_sort_indices(v, order=order, null_placement=np_))


def test_grouper_random(ctx):
    # TestGrouper::ValidateConsume (row/grouper_test.cc:736-760): Take(uniques, ids) == keys, uniques prefix-stable
    for kt in (pa.int64(), pa.int32(), pa.uint8(), pa.float64()):
        g = bc.Grouper([kt], ctx)
        og = ora.Grouper([kt])
        prev = None
        for b in range(4):
            keys = random_array(kt, 3000, 0.05, SEED + b, lo=0, hi=40 * (b + 1))
            ids = g.consume(dev(keys, ctx)).to_arrow()
            assert_equal(ids, og.consume(keys), f"{kt} batch {b}")   # first-occurrence order == oracle
            uniq = g.get_uniques()[0].to_arrow()
            assert g.num_groups == og.num_groups == len(uniq)
            assert pc.take(uniq, ids).equals(keys)
            if prev is not None:
                assert uniq.slice(0, len(prev)).equals(prev)
            prev = uniq
        look = random_array(kt, 500, 0.05, SEED + 99, lo=0, hi=400)
        assert_equal(g.lookup(dev(look, ctx)).to_arrow(), og.lookup(look))
    # multi-column keys with nulls
    g = bc.Grouper([pa.int32(), pa.int16()], ctx)
    og = ora.Grouper([pa.int32(), pa.int16()])
    k0, k1 = random_array(pa.int32(), 5000, 0.1, SEED, lo=0, hi=9), random_array(pa.int16(), 5000, 0.1, SEED + 1, lo=-3, hi=3)
    assert_equal(g.consume([dev(k0, ctx), dev(k1, ctx)]).to_arrow(), og.consume([k0, k1]))
    for got, want in zip(g.get_uniques(), og.get_uniques()):
        assert_equal(got.to_arrow(), want)
    # high-cardinality batch forces table growth
    g = bc.Grouper([pa.int64()], ctx)
    keys = pa.array(np.random.default_rng(SEED).permutation(300000), pa.int64())
    ids = g.consume(dev(keys, ctx)).to_arrow()
    assert ids.to_numpy().tolist() == list(range(300000)) and g.num_groups == 300000


@pytest.mark.parametrize("vt", [pa.int64(), pa.int32(), pa.uint16(), pa.uint64(), pa.float64(), pa.float32()], ids=str)
def test_hash_aggregates_random(ctx, vt):
    n = 20000
    keys = random_array(pa.int64(), n, 0.05, SEED, lo=0, hi=300)
    vals = random_array(vt, n, 0.1, SEED + 3, lo=-100 if not pa.types.is_unsigned_integer(vt) else 0, hi=100)
    aggs = [("hash_sum", None), ("hash_count", None), ("hash_count", {"mode": "only_null"}), ("hash_count", {"mode": "all"}),
            ("hash_mean", None), ("hash_min", None), ("hash_max", None), ("hash_sum", {"skip_nulls": False}),
            ("hash_sum", {"min_count": 60})]
    dk, dvv = dev(keys, ctx), dev(vals, ctx)
    uniq, outs = bc.group_by([dk], [(fn, dvv, o) for fn, o in aggs] + [("hash_count_all", None, None)])
    ouniq, oouts = ora.group_by([keys], [(fn, vals, o) for fn, o in aggs] + [("hash_count_all", None, None)])
    assert_equal(uniq[0].to_arrow(), ouniq[0])
    for (fn, o), got, want in zip(aggs + [("hash_count_all", None)], outs, oouts):
        got = got.to_arrow()
        if pa.types.is_floating(got.type) and fn in ("hash_sum", "hash_mean"):
            assert got.is_valid().equals(want.is_valid())
            np.testing.assert_allclose(got.fill_null(0).to_numpy(), want.fill_null(0).to_numpy(), rtol=1e-9, atol=1e-9)
        else:
            assert_equal(got, want, f"{vt} {fn} {o}")
    # against the reference engine (Acero), sorted by key as its own tests do
    import pyarrow.acero  # noqa: F401
    ref = pa.table({"k": keys, "v": vals}).group_by("k", use_threads=False).aggregate([("v", "sum"), ("v", "count")]).sort_by("k")
    mine = pa.table({"k": uniq[0].to_arrow(), "v_sum": outs[0].to_arrow(), "v_count": outs[1].to_arrow()}).sort_by("k")
    assert mine["k"].combine_chunks().equals(ref["k"].combine_chunks())
    assert mine["v_count"].combine_chunks().equals(ref["v_count"].combine_chunks())
    if pa.types.is_integer(vt):
        assert mine["v_sum"].combine_chunks().equals(ref["v_sum"].combine_chunks())
    # fused group-by == grouper + aggregators
    f = bc.GroupBySumCount(pa.int64(), vt, ctx=ctx)
    half = n // 2
    f.consume(dk.slice(0, half), dvv.slice(0, half))
    f.consume(dk.slice(half), dvv.slice(half))
    fk, fs, fc = [x.to_arrow() for x in f.finalize()]
    fused = pa.table({"k": fk, "v_sum": fs, "v_count": fc}).sort_by("k")
    assert fused["k"].combine_chunks().equals(ref["k"].combine_chunks())
    assert fused["v_count"].combine_chunks().equals(ref["v_count"].combine_chunks())
    if pa.types.is_integer(vt):
        assert fused["v_sum"].combine_chunks().equals(ref["v_sum"].combine_chunks())
    else:
        np.testing.assert_allclose(fused["v_sum"].combine_chunks().fill_null(0).to_numpy(),
                                   ref["v_sum"].combine_chunks().fill_null(0).to_numpy(), rtol=1e-9, atol=1e-9)


def test_hash_aggregate_merge(ctx):
    # GroupByNode::Merge (acero/groupby_aggregate_node.cc:255-298): consume uniques -> transposition -> merge
    n = 8000
    keys = random_array(pa.int64(), n, 0.05, SEED, lo=0, hi=200)
    vals = random_array(pa.int64(), n, 0.1, SEED + 3, lo=-100, hi=100)
    half = n // 2
    states = []
    for sl in (slice(0, half), slice(half, n)):
        k, v = keys.slice(sl.start, sl.stop - sl.start), vals.slice(sl.start, sl.stop - sl.start)
        g = bc.Grouper([pa.int64()], ctx)
        ids = g.consume(dev(k, ctx))
        aggs = [bc.HashAggregator(fn, pa.int64(), ctx=ctx) for fn in ("hash_sum", "hash_count", "hash_min", "hash_max")]
        for a in aggs:
            a.resize(g.num_groups)
            a.consume(dev(v, ctx), ids)
        states.append((g, aggs))
    (g0, a0), (g1, a1) = states
    mapping = g0.consume(g1.get_uniques())
    for x, y in zip(a0, a1):
        x.resize(g0.num_groups)
        x.merge(y, mapping)
    uniq = g0.get_uniques()[0].to_arrow()
    ouniq, oouts = ora.group_by([keys], [(fn, vals, None) for fn in ("hash_sum", "hash_count", "hash_min", "hash_max")])
    assert_equal(uniq, ouniq[0])
    for a, want in zip(a0, oouts):
        assert_equal(a.finalize().to_arrow(), want)


def test_large_filter_take_sort_properties(ctx):
    """BASELINE-sized shapes are checked through size-independent properties (16M rows here so the
    default suite stays in minutes; bench.py runs the 1B-row configuration)."""
    n = 1 << 24
    rng = np.random.default_rng(SEED)
    vals = pa.array(rng.integers(-100, 100, n, dtype=np.int64), pa.int64(), mask=rng.random(n) < 0.1)
    mask = pa.array(rng.random(n) < 0.5)
    dv, dm = dev(vals, ctx), dev(mask, ctx)
    out = bc.filter(dv, dm)
    assert len(out) == int(np.count_nonzero(mask.to_numpy(zero_copy_only=False)))
    assert out.to_arrow().equals(pc.filter(vals, mask))
    # take(iota) is the identity; take(reversed) reverses
    iota = dev(pa.array(np.arange(n, dtype=np.int64)), ctx)
    assert bc.take(dv, iota).to_arrow().equals(vals)
    # sort: output is a permutation, keys non-decreasing along it, nulls last, ties by index
    keys = pa.array(rng.integers(-2**62, 2**62, n, dtype=np.int64), pa.int64(), mask=rng.random(n) < 0.1)
    idx = bc.array_sort_indices(dev(keys, ctx)).to_arrow().to_numpy()
    assert np.array_equal(np.sort(idx), np.arange(n, dtype=np.uint64))
    kv = keys.to_numpy(zero_copy_only=False)
    valid = ora.validity(keys)
    nv = int(valid.sum())
    assert valid[idx[:nv]].all() and not valid[idx[nv:]].any()
    assert np.all(np.diff(idx[nv:].astype(np.int64)) > 0)
    sk = np.frombuffer(keys.buffers()[1], dtype=np.int64)[idx[:nv].astype(np.int64)]
    assert np.all(sk[1:] >= sk[:-1])
    del kv


# ---------------------------------------------------------------- unique / value_counts / dictionary_encode
def test_kat_vector_hash(ctx):
    from tests.test_oracle import VECTOR_HASH_KAT
    for t in (pa.int8(), pa.uint16(), pa.int32(), pa.int64(), pa.float32(), pa.float64()):
        for vals, want in VECTOR_HASH_KAT["unique"]:
            assert_equal(bc.unique(dev(pa.array(vals, t), ctx)).to_arrow(), pa.array(want, t))
        for vals, uniq, counts in VECTOR_HASH_KAT["value_counts"]:
            v, c = bc.value_counts(dev(pa.array(vals, t), ctx))
            assert_equal(v.to_arrow(), pa.array(uniq, t))
            assert_equal(c.to_arrow(), pa.array(counts, pa.int64()))
        for vals, dictionary, idx in VECTOR_HASH_KAT["dictionary_encode"]:
            got = bc.dictionary_encode(dev(pa.array(vals, t), ctx)).to_arrow()
            assert got.equals(pa.DictionaryArray.from_arrays(pa.array(idx, pa.int32()), pa.array(dictionary, t)))
        for empty in (pa.array([], t), pa.array([None, None], t)):
            assert_equal(bc.unique(dev(empty, ctx)).to_arrow(), pc.unique(empty))
            assert bc.dictionary_encode(dev(empty, ctx)).to_arrow().equals(pc.dictionary_encode(empty))
            assert bc.dictionary_encode(dev(empty, ctx), "encode").to_arrow().equals(pc.dictionary_encode(empty, "encode"))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_vector_hash_random(ctx, t):
    for n, hi, null_p in ((1, 3, 0.0), (300, 7, 0.2), (70000, 100, 0.05), (70000, 5000, 0.0), (3000, 3000, 1.0),
                          (300000, 120, 0.01)):
        a = random_array(t, n, null_p, SEED + n, lo=0, hi=hi, offset=3)
        d = dev(a, ctx)
        assert_equal(bc.unique(d).to_arrow(), pc.unique(a), f"unique {t} {n}")
        assert_equal(bc.unique(d).to_arrow(), ora.unique(a))
        v, c = bc.value_counts(d)
        assert bc.value_counts_to_struct(v, c).equals(pc.value_counts(a)), f"value_counts {t} {n}"
        for mode in ("mask", "encode"):
            got = bc.dictionary_encode(d, mode).to_arrow()
            assert got.equals(pc.dictionary_encode(a, null_encoding=mode)), f"dictionary_encode {t} {n} {mode}"
            assert got.equals(ora.dictionary_encode(a, mode))
    assert bc.call_function("unique", [dev(pa.array([3, 3, 1], t), ctx)]).to_arrow().equals(pa.array([3, 1], t))
    assert bc.unique(dev(pa.array(["a", "b", "a"]), ctx)).to_arrow().equals(pa.array(["a", "b"]))  # r2: strings too
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.unique(dev(pa.array([True, False]), ctx))


# ---------------------------------------------------------------- ungrouped sum / mean / min_max / count
@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_scalar_aggregates(ctx, t):
    acc = ora._agg_acc(t)[0]
    assert bc.sum(dev(pa.array([0, 1, 2, 3, 4, 5], t), ctx)) == pa.scalar(15, acc)          # TestNumericSumKernel.SimpleSum
    assert bc.sum(dev(pa.array([0, None, 2, 3, None, 5], t), ctx)) == pa.scalar(10, acc)
    assert bc.mean(dev(pa.array([1, 2, 3, 4, 5, 6, 7, 8], t), ctx)) == pa.scalar(4.5, pa.float64())
    assert bc.min_max(dev(pa.array([5, None, 2, 3, 4], t), ctx)).as_py() == {"min": 2, "max": 5}
    # offset 5: values not 16-byte aligned (scalar loads); offsets 0 / 16: the vectorised path, 16 with a bit offset
    for n, null_p, off in ((0, 0.0, 5), (1, 0.0, 5), (1000, 0.0, 5), (70001, 0.1, 5), (5000, 1.0, 5), (1 << 21, 0.3, 5),
                           (70001, 0.1, 0), (1 << 21, 0.3, 16), (1000, 0.0, 16), (37, 0.5, 0)):
        a = random_array(t, n, null_p, SEED + n, offset=off)
        d = dev(a, ctx)
        for skip in (True, False):
            for mc in (0, 1, 3):
                want, got = pc.sum(a, skip_nulls=skip, min_count=mc), bc.sum(d, skip, mc)
                if pa.types.is_floating(t) and want.is_valid:
                    # float sums: the reference adds pairwise, the device in a fixed tree order (tolerance 1e-12 relative
                    # to the sum of magnitudes)
                    scale = float(np.abs(ora.values(a)[ora.validity(a)].astype(np.float64)).sum()) or 1.0
                    assert got.is_valid and abs(got.as_py() - want.as_py()) <= 1e-12 * scale
                else:
                    assert got == want, (t, n, null_p, skip, mc)  # integer sums are bit-exact (wrapping)
                wm, gm = pc.mean(a, skip_nulls=skip, min_count=mc), bc.mean(d, skip, mc)
                assert gm.is_valid == wm.is_valid
                if wm.is_valid:
                    scale = float(np.abs(ora.values(a)[ora.validity(a)].astype(np.float64)).sum()) / max(1, len(a) - a.null_count)
                    assert np.isnan(wm.as_py()) and np.isnan(gm.as_py()) or abs(gm.as_py() - wm.as_py()) <= 1e-12 * (scale or 1.0)
                assert bc.min_max(d, skip, mc) == pc.min_max(a, skip_nulls=skip, min_count=mc), (t, n, null_p, skip, mc)
                assert bc.min_max(d, skip, mc) == ora.scalar_min_max(a, skip, mc)
        for mode in ("only_valid", "only_null", "all"):
            assert bc.count(d, mode) == pc.count(a, mode=mode)
            assert bc.count(d.slice(3, max(0, n - 7)), mode) == pc.count(a.slice(3, max(0, n - 7)), mode=mode)
    if pa.types.is_floating(t):
        nan = pa.array([np.nan, 1.0, None, -2.0, np.nan], t)
        assert bc.min_max(dev(nan, ctx)) == pc.min_max(nan)
        assert np.isnan(bc.min_max(dev(pa.array([np.nan, np.nan], t), ctx))["max"].as_py())
    assert bc.call_function("sum", [dev(pa.array([1, 2], t), ctx)]) == pa.scalar(3, acc)


@pytest.mark.parametrize("vt", [pa.int64(), pa.int16(), pa.uint32(), pa.float64(), pa.float32()], ids=str)
@pytest.mark.parametrize("groups", [1, 7, 2048, 2049])
def test_hash_aggregates_few_groups_large(ctx, vt, groups):
    """<= 2048 groups take the shared-memory privatised consume kernel (hot groups would serialise their
    global atomics); 2049 the plain one.  Oracle = the reference engine (Acero group_by), sorted by key."""
    import pyarrow.acero  # noqa: F401
    n = 1 << 20
    keys = random_array(pa.int64(), n, 0.02, SEED + groups, lo=0, hi=groups - 1)
    vals = random_array(vt, n, 0.1, SEED + 3, lo=-100 if not pa.types.is_unsigned_integer(vt) else 0, hi=100)
    fns = ["sum", "count", "mean", "min", "max"]
    uniq, outs = bc.group_by([dev(keys, ctx)], [("hash_" + f, dev(vals, ctx), None) for f in fns] + [("hash_count_all", None, None)])
    ref = pa.table({"k": keys, "v": vals}).group_by("k", use_threads=False).aggregate([("v", f) for f in fns] + [([], "count_all")]).sort_by("k")
    mine = pa.table({"k": uniq[0].to_arrow(), **{"v_" + f: o.to_arrow() for f, o in zip(fns, outs)}, "count_all": outs[-1].to_arrow()}).sort_by("k")
    assert mine["k"].combine_chunks().equals(ref["k"].combine_chunks())
    for f in fns + ["count_all"]:
        name = f if f == "count_all" else "v_" + f
        got, want = mine[name].combine_chunks(), ref[name].combine_chunks()
        if pa.types.is_floating(got.type) and f in ("sum", "mean"):
            assert got.is_valid().equals(want.is_valid())
            np.testing.assert_allclose(got.fill_null(0).to_numpy(), want.fill_null(0).to_numpy(), rtol=1e-9, atol=1e-6)
        else:
            assert got.equals(want), (vt, groups, f)


@pytest.mark.parametrize("t", [pa.int64(), pa.float32(), pa.uint16()], ids=str)
def test_sort_payload_equals_sort_indices_then_take(ctx, t):
    """b2_sort_payload: the payload rides along the radix passes; result = take(payload, sort_indices(values))"""
    for n, null_p in ((0, 0.0), (1, 0.0), (5000, 0.1), (200_003, 0.3)):
        vals = random_array(t, n, null_p, SEED + n, lo=0, hi=50, offset=3)     # many ties: stability matters
        payload = pa.array(np.random.default_rng(SEED).integers(0, 2**32, n, dtype=np.uint32))
        for order in ("ascending", "descending"):
            for placement in ("at_end", "at_start"):
                got = bc.sort_payload(dev(vals, ctx), dev(payload, ctx), order, placement).to_arrow()
                want = pc.take(payload, pc.array_sort_indices(vals, order=order, null_placement=placement)).cast(pa.uint64())
                assert got.equals(want), f"{t} n={n} {order} {placement}"


@pytest.mark.parametrize("vt", [pa.int32(), pa.uint16(), pa.float32(), pa.float64(), pa.int64()], ids=str)
def test_group_by_mean_rides_the_fused_state(ctx, vt):
    # hash_mean = sum / count of the fused table (GroupedMeanImpl::DoMean, hash_aggregate_numeric.cc:398-402); int64
    # columns stay on the unfused kernel -- either way the answer is the reference engine's
    n = 50000
    keys = random_array(pa.int32(), n, 0.05, SEED, lo=-50, hi=400)
    vals = random_array(vt, n, 0.2, SEED + 5, lo=0 if pa.types.is_unsigned_integer(vt) else -1000, hi=1000)
    dk, dvv = dev(keys, ctx), dev(vals, ctx)
    uniq, (mean, cnt, total) = bc.group_by([dk], [("hash_mean", dvv, None), ("hash_count", dvv, None), ("hash_sum", dvv, None)])
    mine = pa.table({"k": uniq[0].to_arrow(), "m": mean.to_arrow(), "c": cnt.to_arrow(), "s": total.to_arrow()}).sort_by("k")
    import pyarrow.acero  # noqa: F401
    ref = pa.table({"k": keys, "v": vals}).group_by("k", use_threads=False).aggregate([("v", "mean"), ("v", "count"), ("v", "sum")]).sort_by("k")
    assert mine["k"].combine_chunks().equals(ref["k"].combine_chunks())
    assert mine["c"].combine_chunks().equals(ref["v_count"].combine_chunks())
    got, want = mine["m"].combine_chunks(), ref["v_mean"].combine_chunks()
    assert got.type == pa.float64() and got.is_valid().equals(want.is_valid())
    np.testing.assert_allclose(got.fill_null(0).to_numpy(), want.fill_null(0).to_numpy(), rtol=1e-12, atol=1e-12)
    # and the unfused path agrees
    _, (mean2,) = bc.group_by([dk], [("hash_mean", dvv, None)], fused=False)
    uniq2, _ = bc.group_by([dk], [("hash_count", dvv, None)], fused=False)
    m2 = pa.table({"k": uniq2[0].to_arrow(), "m": mean2.to_arrow()}).sort_by("k")["m"].combine_chunks()
    np.testing.assert_allclose(m2.fill_null(0).to_numpy(), want.fill_null(0).to_numpy(), rtol=1e-12, atol=1e-12)
