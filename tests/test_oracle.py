"""Pins the oracle (oracle/arrow_oracle.py): (a) the reference's own known-answer vectors
(tests/golden/kat.json) and (b) the reference binary itself (pyarrow 24.0.0's
libarrow_compute) on seeded random inputs.  CPU only."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import arrow_oracle as ora
from tests.util import INT_TYPES, NUMERIC_TYPES, SEED, TYPE_BY_NAME, assert_equal, equal_nan, from_json, kat, random_array

KAT = kat()


# ---------------------------------------------------------------- (a) known answers
@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_kat_filter(t):
    for c in KAT["filter_numeric_basics"]["cases"]:
        v, m = from_json(t, c["values"]), from_json(pa.bool_(), c["filter"])
        assert_equal(ora.filter(v, m, "emit_null"), from_json(t, c["emit_null"]), f"emit_null {c}")
        assert_equal(ora.filter(v, m, "drop"), from_json(t, c["drop"]), f"drop {c}")
    with pytest.raises(pa.ArrowInvalid):
        ora.filter(from_json(t, [7, 8, 9]), from_json(pa.bool_(), []), "drop")


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_kat_take(t):
    for c in KAT["take_numeric_basics"]["cases"]:
        for it in (pa.int8(), pa.uint32(), pa.int64()):
            assert_equal(ora.take(from_json(t, c["values"]), from_json(it, c["indices"])), from_json(t, c["expected"]))
    for c in KAT["take_numeric_basics"]["index_errors"]:
        with pytest.raises(pa.ArrowIndexError):
            ora.take(from_json(t, c["values"]), from_json(pa.int8(), c["indices"]))


@pytest.mark.parametrize("t", INT_TYPES, ids=str)
def test_kat_sort_integral(t):
    for c in KAT["sort_integral"]["cases"]:
        got = ora.sort_indices(from_json(t, c["values"]), c["order"], c["null_placement"])
        assert_equal(got, pa.array(c["expected"], pa.uint64()), str(c))
    if t == pa.int64():
        for c in KAT["sort_integral"]["int64_cases"]:
            got = ora.sort_indices(from_json(t, c["values"]), c["order"], c["null_placement"])
            assert_equal(got, pa.array(c["expected"], pa.uint64()), str(c))


@pytest.mark.parametrize("t", [pa.float32(), pa.float64()], ids=str)
def test_kat_sort_real(t):
    for c in KAT["sort_real"]["cases"]:
        got = ora.sort_indices(from_json(t, c["values"]), c["order"], c["null_placement"])
        assert_equal(got, pa.array(c["expected"], pa.uint64()), str(c))


def test_kat_cast():
    for c in KAT["cast_float_to_float"]["cases"]:
        assert_equal(ora.cast_array(from_json(TYPE_BY_NAME[c["from"]], c["values"]), TYPE_BY_NAME[c["to"]]),
                     from_json(TYPE_BY_NAME[c["to"]], c["expected"]))
    for c in KAT["cast_int_to_float_bounds"]["ok"]:
        ora.cast_array(from_json(TYPE_BY_NAME[c["from"]], c["values"]), TYPE_BY_NAME[c["to"]])
    for c in KAT["cast_int_to_float_bounds"]["fails"]:
        with pytest.raises(pa.ArrowInvalid):
            ora.cast_array(from_json(TYPE_BY_NAME[c["from"]], c["values"]), TYPE_BY_NAME[c["to"]])
    c = KAT["cast_overflow_in_null_slot"]
    v = pa.array(c["values"], TYPE_BY_NAME[c["from"]], mask=~np.array(c["validity"], dtype=bool))
    assert_equal(ora.cast_array(v, TYPE_BY_NAME[c["to"]]), from_json(TYPE_BY_NAME[c["to"]], c["expected"]))


def _sorted_by_key(keys: pa.Array, cols):
    order = pc.sort_indices(keys, null_placement="at_end")
    return [pc.take(c, order) for c in [keys] + list(cols)]


def test_kat_group_by():
    c = KAT["group_by_count_only"]
    arg = pa.array([r[0] for r in c["rows"]], pa.float64())
    key = pa.array([r[1] for r in c["rows"]], pa.int64())
    for mode in ("only_valid", "only_null", "all"):
        uniq, (cnt,) = ora.group_by([key], [("hash_count", arg, {"mode": mode})])
        k, v = _sorted_by_key(uniq[0], [cnt])
        assert k.to_pylist() == [r[0] for r in c[mode]] and v.to_pylist() == [r[1] for r in c[mode]]
    c = KAT["group_by_sum_only"]
    arg = pa.array([r[0] for r in c["rows"]], pa.float64())
    key = pa.array([r[1] for r in c["rows"]], pa.int64())
    uniq, (s,) = ora.group_by([key], [("hash_sum", arg, None)])
    k, v = _sorted_by_key(uniq[0], [s])
    assert k.to_pylist() == [r[0] for r in c["expected"]] and v.to_pylist() == [r[1] for r in c["expected"]]
    c = KAT["grouper_int64"]
    g = ora.Grouper([pa.int64()])
    assert g.consume(pa.array(c["keys"], pa.int64())).to_pylist() == c["ids"]


# ---------------------------------------------------------------- (b) the reference binary
NULL_PROBS = [0.0, 0.1, 0.999]


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("null_p", NULL_PROBS)
def test_filter_vs_reference(t, null_p):
    for i, (true_p, mask_null, off) in enumerate([(0.5, 0.0, 0), (0.1, 0.05, 3), (0.999, 0.5, 2)]):
        v = random_array(t, 1024, null_p, SEED + i, offset=off)
        m = random_array(pa.bool_(), 1024, mask_null, SEED + 77 + i, hi=true_p, offset=(off * 5) % 7)
        for ns in ("drop", "emit_null"):
            assert_equal(ora.filter(v, m, ns), pc.filter(v, m, null_selection_behavior=ns), f"{t} {ns}")


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("it", [pa.int8(), pa.uint16(), pa.int32(), pa.uint32(), pa.int64(), pa.uint64()], ids=str)
def test_take_vs_reference(t, it):
    for null_p in (0.0, 0.05, 0.95):
        v = random_array(t, 1025, null_p, SEED, offset=1)
        hi = min(1024, np.iinfo(it.to_pandas_dtype()).max)
        idx = random_array(it, 257, null_p, SEED + 5, lo=0, hi=hi, offset=3)
        assert_equal(ora.take(v, idx), pc.take(v, idx), f"{t} {it} {null_p}")


@pytest.mark.parametrize("src", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("dst", NUMERIC_TYPES, ids=str)
def test_cast_vs_reference(src, dst):
    if src == dst:
        return
    v = random_array(src, 500, 0.1, SEED, lo=0, hi=100, offset=1)  # in range for every target
    assert_equal(ora.cast_array(v, dst, safe=False), pc.cast(v, dst, safe=False), f"{src}->{dst}")
    if pa.types.is_floating(src) and pa.types.is_integer(dst):
        w = pc.round(v)
        assert_equal(ora.cast_array(w, dst, safe=True), pc.cast(w, dst, safe=True))
        with pytest.raises(pa.ArrowInvalid):
            ora.cast_array(pa.array([1.5], src), dst, safe=True)
        with pytest.raises(pa.ArrowInvalid):
            pc.cast(pa.array([1.5], src), dst, safe=True)
    else:
        assert_equal(ora.cast_array(v, dst, safe=True), pc.cast(v, dst, safe=True))
    # wide-range input: wraps when unsafe, errors (both sides) when safe and out of range
    wide = random_array(src, 300, 0.1, SEED + 1)
    if pa.types.is_integer(src) and pa.types.is_integer(dst):
        assert_equal(ora.cast_array(wide, dst, safe=False), pc.cast(wide, dst, safe=False))
        try:
            want = pc.cast(wide, dst, safe=True)
        except pa.ArrowInvalid as e:
            with pytest.raises(pa.ArrowInvalid) as ei:
                ora.cast_array(wide, dst, safe=True)
            assert str(ei.value) == str(e)
        else:
            assert_equal(ora.cast_array(wide, dst, safe=True), want)


ARITH = ["add", "subtract", "multiply", "add_checked", "subtract_checked", "multiply_checked", "divide", "divide_checked"]


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("op", ARITH)
def test_arithmetic_vs_reference(t, op):
    small = dict(lo=1, hi=11) if pa.types.is_integer(t) else dict(lo=-100, hi=100)
    cases = [(random_array(t, 300, 0.1, SEED, offset=1, **small), random_array(t, 300, 0.1, SEED + 1, offset=2, **small))]
    if not op.endswith("_checked") and "divide" not in op:
        cases.append((random_array(t, 300, 0.1, SEED + 2), random_array(t, 300, 0.0, SEED + 3)))  # wraps
    for a, b in cases:
        try:
            want = getattr(pc, op)(a, b)
        except pa.ArrowInvalid as e:  # e.g. unsigned subtract_checked underflow: same error both sides
            with pytest.raises(pa.ArrowInvalid) as ei:
                ora.arithmetic(op, a, b)
            assert str(ei.value) == str(e)
            continue
        got = ora.arithmetic(op, a, b)
        assert equal_nan(got, want), f"{t} {op}"
        assert equal_nan(ora.arithmetic(op, a, b[0]), getattr(pc, op)(a, b[0]))
        assert equal_nan(ora.arithmetic(op, a[0], b), getattr(pc, op)(a[0], b))


def test_arithmetic_errors_vs_reference():
    for t in INT_TYPES:
        info = np.iinfo(t.to_pandas_dtype())
        a = pa.array([info.max, 1], t)
        for op in ("add_checked", "multiply_checked"):
            with pytest.raises(pa.ArrowInvalid, match="overflow"):
                getattr(pc, op)(a, pa.array([2, 1], t))
            with pytest.raises(pa.ArrowInvalid, match="overflow"):
                ora.arithmetic(op, a, pa.array([2, 1], t))
        for op in ("divide", "divide_checked"):
            with pytest.raises(pa.ArrowInvalid, match="divide by zero"):
                getattr(pc, op)(a, pa.array([0, 1], t))
            with pytest.raises(pa.ArrowInvalid, match="divide by zero"):
                ora.arithmetic(op, a, pa.array([0, 1], t))
        # errors under nulls are ignored
        z = pa.array([0, 1], t, mask=np.array([True, False]))
        assert_equal(ora.arithmetic("divide", a, z), pc.divide(a, z))
    assert_equal(ora.arithmetic("divide", pa.array([-128], pa.int8()), pa.array([-1], pa.int8())),
                 pc.divide(pa.array([-128], pa.int8()), pa.array([-1], pa.int8())))


def test_mixed_type_dispatch_vs_reference():
    pairs = [(pa.int8(), pa.uint8()), (pa.int32(), pa.uint32()), (pa.uint64(), pa.int8()), (pa.int64(), pa.float32()),
             (pa.uint16(), pa.float64()), (pa.int16(), pa.int64())]
    for ta, tb in pairs:
        a, b = random_array(ta, 100, 0.1, SEED, lo=0, hi=50), random_array(tb, 100, 0.1, SEED + 9, lo=0, hi=50)
        assert_equal(ora.arithmetic("add", a, b), pc.add(a, b), f"{ta}+{tb}")
        assert_equal(ora.compare("less", a, b), pc.less(a, b), f"{ta}<{tb}")


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_compare_vs_reference(t):
    a = random_array(t, 777, 0.1, SEED, lo=0, hi=20, offset=3)
    b = random_array(t, 777, 0.1, SEED + 1, lo=0, hi=20, offset=5)
    for op in ("equal", "not_equal", "greater", "greater_equal", "less", "less_equal"):
        assert_equal(ora.compare(op, a, b), getattr(pc, op)(a, b), f"{t} {op}")
        assert_equal(ora.compare(op, a, b[1]), getattr(pc, op)(a, b[1]))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
@pytest.mark.parametrize("null_p", [0.0, 0.1, 0.5, 1.0])
def test_sort_vs_reference(t, null_p):
    for i, rng in enumerate([dict(), dict(lo=0, hi=10)]):
        v = random_array(t, 1000, null_p, SEED + i, offset=i, **rng)
        if pa.types.is_floating(t):
            a = v.to_numpy(zero_copy_only=False).copy()
            a[::17] = np.nan
            a[5::31] = -0.0
            a[7::31] = 0.0
            v = pa.array(a, t, mask=~ora.validity(v))
        for order in ("ascending", "descending"):
            for np_ in ("at_end", "at_start"):
                want = pc.array_sort_indices(v, order=order, null_placement=np_)
                assert_equal(ora.sort_indices(v, order, np_), want, f"{t} {order} {np_}")


def test_grouper_and_aggregates_vs_reference():
    import pyarrow.acero  # noqa: F401
    n = 5000
    keys = random_array(pa.int64(), n, 0.05, SEED, lo=0, hi=300)
    for vt in (pa.int64(), pa.int32(), pa.uint16(), pa.float64(), pa.float32()):
        vals = random_array(vt, n, 0.1, SEED + 3, lo=-100 if not pa.types.is_unsigned_integer(vt) else 0, hi=100)
        tbl = pa.table({"k": keys, "v": vals})
        want = tbl.group_by("k", use_threads=False).aggregate(
            [("v", "sum"), ("v", "count"), ("v", "mean"), ("v", "min"), ("v", "max"), ([], "count_all")])
        uniq, outs = ora.group_by([keys], [("hash_sum", vals, None), ("hash_count", vals, None), ("hash_mean", vals, None),
                                           ("hash_min", vals, None), ("hash_max", vals, None), ("hash_count_all", None, None)])
        got = pa.table({"k": uniq[0], "v_sum": outs[0], "v_count": outs[1], "v_mean": outs[2], "v_min": outs[3],
                        "v_max": outs[4], "count_all": outs[5]}).sort_by("k")
        want = want.sort_by("k").select(got.column_names)
        for name in got.column_names:
            g, w = got[name].combine_chunks(), want[name].combine_chunks()
            if pa.types.is_floating(g.type) and name in ("v_sum", "v_mean"):
                assert g.is_valid().equals(w.is_valid())
                np.testing.assert_allclose(g.fill_null(0).to_numpy(), w.fill_null(0).to_numpy(), rtol=1e-9)
            else:
                assert_equal(g, w, f"{vt} {name}")
    # ids: first-occurrence order == the reference Grouper on this (single-batch) input
    g = ora.Grouper([pa.int64()])
    ids = g.consume(keys)
    assert pc.take(g.get_uniques()[0], ids).equals(keys)


STRING_TYPES = [pa.string(), pa.large_string(), pa.binary(), pa.large_binary()]


@pytest.mark.parametrize("t", STRING_TYPES, ids=str)
def test_binary_filter_take_vs_reference(t):
    # string KATs of vector_selection_test.cc:670-698 / 1661-1676 shape + random (0-32 byte strings)
    v = pa.array(["a", "b", "c"], pa.string()).cast(t)
    assert_equal(ora.filter(v, pa.array([False, True, False])), v.slice(1, 1))
    m = pa.array([None, True, False])
    assert_equal(ora.filter(v, m, "emit_null"), pc.filter(v, m, null_selection_behavior="emit_null"))
    assert_equal(ora.take(v, pa.array([2, None, 0, 0], pa.int8())), pc.take(v, pa.array([2, None, 0, 0], pa.int8())))
    for null_p in (0.0, 0.1, 0.9):
        vals = random_array(t, 1500, null_p, SEED, offset=3)
        mask = random_array(pa.bool_(), 1500, 0.05, SEED + 1, hi=0.5, offset=2)
        for ns in ("drop", "emit_null"):
            assert_equal(ora.filter(vals, mask, ns), pc.filter(vals, mask, null_selection_behavior=ns), f"{t} {ns}")
        idx = random_array(pa.int32(), 700, null_p, SEED + 2, lo=0, hi=1499, offset=1)
        assert_equal(ora.take(vals, idx), pc.take(vals, idx))
    with pytest.raises(pa.ArrowIndexError):
        ora.take(v, pa.array([0, 3], pa.int32()))


def test_dictionary_filter_take_vs_reference():
    # TakeDictionary / FilterDictionary (vector_selection_test.cc:1678-1683): indices move, dictionary passes through
    d = pa.DictionaryArray.from_arrays(random_array(pa.int32(), 2000, 0.1, SEED, lo=0, hi=9),
                                       pa.array([f"v{i}" for i in range(10)]))
    mask = random_array(pa.bool_(), 2000, 0.05, SEED + 1, hi=0.5)
    for ns in ("drop", "emit_null"):
        assert_equal(ora.filter(d, mask, ns), pc.filter(d, mask, null_selection_behavior=ns))
    idx = random_array(pa.int64(), 500, 0.1, SEED + 2, lo=0, hi=1999)
    assert_equal(ora.take(d, idx), pc.take(d, idx))


# KATs transcribed from the reference's own tests (compute/kernels/vector_hash_test.cc):
# TestHashKernelPrimitive Unique :177-180, ValueCounts :209-212, DictEncode :242-244.
VECTOR_HASH_KAT = {
    "unique": [([2, None, 2, 1], [2, None, 1]), ([None, None, 3, 1], [None, 3, 1])],
    "value_counts": [([2, None, 2, 1, 2, 3, None], [2, None, 1, 3], [3, 2, 1, 1])],
    "dictionary_encode": [([2, None, 2, 1, 2, 3], [2, 1, 3], [0, None, 0, 1, 0, 2])],
}


def test_kat_vector_hash():
    for t in (pa.int8(), pa.uint16(), pa.int32(), pa.int64(), pa.float32(), pa.float64()):
        for vals, want in VECTOR_HASH_KAT["unique"]:
            assert_equal(ora.unique(pa.array(vals, t)), pa.array(want, t))
        for vals, uniq, counts in VECTOR_HASH_KAT["value_counts"]:
            assert ora.value_counts(pa.array(vals, t)).equals(pa.StructArray.from_arrays(
                [pa.array(uniq, t), pa.array(counts, pa.int64())], names=["values", "counts"]))
        for vals, dictionary, idx in VECTOR_HASH_KAT["dictionary_encode"]:
            assert ora.dictionary_encode(pa.array(vals, t)).equals(
                pa.DictionaryArray.from_arrays(pa.array(idx, pa.int32()), pa.array(dictionary, t)))
        for empty in (pa.array([], t), pa.array([None, None], t)):
            assert_equal(ora.unique(empty), pc.unique(empty))
            assert ora.dictionary_encode(empty).equals(pc.dictionary_encode(empty))


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_vector_hash_vs_reference(t):
    for n, hi, null_p in ((1, 3, 0.0), (300, 7, 0.2), (5000, 100, 0.05), (5000, 100, 0.0), (2000, 2000, 1.0)):
        a = random_array(t, n, null_p, SEED + n, lo=0, hi=hi, offset=3)
        assert_equal(ora.unique(a), pc.unique(a))
        assert ora.value_counts(a).equals(pc.value_counts(a))
        for mode in ("mask", "encode"):
            assert ora.dictionary_encode(a, mode).equals(pc.dictionary_encode(a, null_encoding=mode)), (t, n, mode)


# Known answers from the reference's tests (compute/kernels/aggregate_test.cc): TestNumericSumKernel
# SimpleSum "[0, 1, 2, 3, 4, 5]" -> 15 and "[0, null, 2, 3, null, 5]" -> 10, empty -> null,
# min_count; TestPrimitiveMinMaxKernel "[5, 1, 2, 3, 4]" -> (1, 5), "[5, null, 2, 3, 4]" -> (2, 5);
# TestMeanKernelNumeric "[1, 2, 3, 4, 5, 6, 7, 8]" -> 4.5.
def test_kat_scalar_aggregates():
    for t in NUMERIC_TYPES:
        acc = ora._agg_acc(t)[0]
        assert ora.scalar_sum(pa.array([0, 1, 2, 3, 4, 5], t)) == pa.scalar(15, acc)
        assert ora.scalar_sum(pa.array([0, None, 2, 3, None, 5], t)) == pa.scalar(10, acc)
        assert ora.scalar_sum(pa.array([], t)) == pa.scalar(None, acc)
        assert ora.scalar_sum(pa.array([], t), min_count=0) == pa.scalar(0, acc)
        assert ora.scalar_sum(pa.array([1, None], t), skip_nulls=False) == pa.scalar(None, acc)
        assert ora.scalar_sum(pa.array([1, None], t), min_count=2) == pa.scalar(None, acc)
        assert ora.scalar_mean(pa.array([1, 2, 3, 4, 5, 6, 7, 8], t)) == pa.scalar(4.5, pa.float64())
        assert ora.scalar_min_max(pa.array([5, 1, 2, 3, 4], t)).as_py() == {"min": 1, "max": 5}
        assert ora.scalar_min_max(pa.array([5, None, 2, 3, 4], t)).as_py() == {"min": 2, "max": 5}
        assert ora.scalar_min_max(pa.array([5, None, 2], t), skip_nulls=False).as_py() == {"min": None, "max": None}
        assert ora.scalar_count(pa.array([5, None, 2], t), "only_null") == pa.scalar(1, pa.int64())


@pytest.mark.parametrize("t", NUMERIC_TYPES, ids=str)
def test_scalar_aggregates_vs_reference(t):
    for n, null_p in ((0, 0.0), (1, 0.0), (1000, 0.0), (70001, 0.1), (5000, 1.0)):
        a = random_array(t, n, null_p, SEED + n, offset=5)
        for skip in (True, False):
            for mc in (0, 1, 3):
                want = pc.sum(a, skip_nulls=skip, min_count=mc)
                got = ora.scalar_sum(a, skip, mc)
                if pa.types.is_floating(t) and want.is_valid:
                    assert got.is_valid and np.isclose(got.as_py(), want.as_py(), rtol=1e-12, atol=0)
                else:
                    assert got == want, (t, n, null_p, skip, mc)
                wm, gm = pc.mean(a, skip_nulls=skip, min_count=mc), ora.scalar_mean(a, skip, mc)
                assert gm.is_valid == wm.is_valid and (not wm.is_valid or np.isclose(gm.as_py(), wm.as_py(), rtol=1e-12, equal_nan=True))
                assert ora.scalar_min_max(a, skip, mc) == pc.min_max(a, skip_nulls=skip, min_count=mc)
        for mode in ("only_valid", "only_null", "all"):
            assert ora.scalar_count(a, mode) == pc.count(a, mode=mode)
    if pa.types.is_floating(t):
        nan = pa.array([np.nan, 1.0, None, -2.0, np.nan], t)
        assert ora.scalar_min_max(nan) == pc.min_max(nan)
        allnan = pa.array([np.nan, np.nan], t)
        assert np.isnan(ora.scalar_min_max(allnan)["min"].as_py()) and np.isnan(pc.min_max(allnan)["min"].as_py())


@pytest.mark.parametrize("t", [pa.string(), pa.large_string(), pa.binary(), pa.large_binary()], ids=str)
def test_string_vector_hash_and_grouper_vs_reference(t):
    """oracle unique / value_counts / dictionary_encode and Grouper over utf8 / binary values against the reference binary
    (vector_hash.cc:782-830; Grouper ids through Take(uniques, ids) == keys, row/grouper_test.cc:736-760)."""
    rng = np.random.default_rng(SEED)
    vocab = [bytes(rng.integers(97, 123, int(rng.integers(0, 12)), dtype=np.uint8)) for _ in range(200)]
    is_bin = pa.types.is_binary(t) or pa.types.is_large_binary(t)
    vals = [None if rng.random() < 0.1 else (vocab[i] if is_bin else vocab[i].decode()) for i in rng.integers(0, 200, 3003)]
    arr = pa.array(vals, t).slice(3)
    assert ora.unique(arr).equals(pc.unique(arr))
    assert ora.value_counts(arr).equals(pc.value_counts(arr))
    for enc in ("mask", "encode"):
        assert ora.dictionary_encode(arr, enc).equals(pc.dictionary_encode(arr, enc))
    other = pa.array(rng.integers(0, 3, len(arr)), pa.int64())
    g = ora.Grouper([t, pa.int64()])
    ids = g.consume([arr, other])
    u = g.get_uniques()
    assert pc.take(u[0], ids).equals(arr) and pc.take(u[1], ids).equals(other)
    ref = pa.table([arr, other], names=["a", "b"]).group_by(["a", "b"], use_threads=False).aggregate([])
    assert g.num_groups == ref.num_rows
    look = g.lookup([pa.array([vals[3], "never-seen" if not is_bin else b"never-seen"], t), pa.array([int(other[0].as_py()), 0], pa.int64())])
    assert look.to_pylist() == [0, None]


@pytest.mark.parametrize("join_type", ["inner", "left outer", "left semi", "left anti", "full outer"])
def test_hash_join_indices_vs_reference(join_type):
    """oracle join pairs against the reference binary's HashJoinNode (pyarrow.Table.join): null keys match nothing,
    every matching pair appears exactly once, unmatched left rows are null-extended for the outer join."""
    rng = np.random.default_rng(SEED + 3)
    n_l, n_r = 700, 500
    lk = [pa.array(rng.integers(0, 40, n_l), pa.int64(), mask=rng.random(n_l) < 0.1),
          pa.array([None if rng.random() < 0.05 else "s" + str(int(x)) for x in rng.integers(0, 3, n_l)], pa.string())]
    rk = [pa.array(rng.integers(0, 45, n_r), pa.int64(), mask=rng.random(n_r) < 0.1),
          pa.array([None if rng.random() < 0.05 else "s" + str(int(x)) for x in rng.integers(0, 3, n_r)], pa.string())]
    l, r = ora.hash_join_indices(lk, rk, join_type)
    lt = pa.table({"a": lk[0], "b": lk[1], "lrow": np.arange(n_l)})
    rt = pa.table({"a": rk[0], "b": rk[1], "rrow": np.arange(n_r)})
    out = lt.join(rt, keys=["a", "b"], join_type=join_type, use_threads=False)
    rrow = out["rrow"].to_pylist() if "rrow" in out.column_names else [None] * out.num_rows
    ref = sorted(zip([(-1 if x is None else x) for x in out["lrow"].to_pylist()], [(-1 if x is None else x) for x in rrow]))
    mine = sorted(zip([(-1 if x is None else x) for x in l.to_pylist()],
                      [(-1 if x is None else x) for x in (r.to_pylist() if r is not None else [None] * len(l))]))
    assert mine == ref
    head = [x for x in l.to_pylist() if x is not None]
    assert head == sorted(head)  # left-row order


@pytest.mark.parametrize("t", [pa.int64(), pa.float64(), pa.uint16()], ids=str)
@pytest.mark.parametrize("order", ["ascending", "descending"])
def test_select_k_vs_reference(t, order):
    """oracle select_k_unstable against the reference binary: same selected VALUES while k stays inside the non-null,
    non-NaN values (ties may be broken differently; the 24.0.0 binary stops at the last such value)."""
    arr = random_array(t, 5000, 0.1, SEED + 21, lo=0, hi=300, offset=2)
    n_values = len(arr) - arr.null_count
    for k in (0, 1, 17, 999, n_values):
        mine = ora.select_k_unstable(arr, k, order)
        ref = pc.select_k_unstable(arr, k, [("x", order)])
        assert len(mine) == len(ref) == k
        assert pc.take(arr, mine).equals(pc.take(arr, ref))
    with pytest.raises(pa.ArrowInvalid):
        ora.select_k_unstable(arr, -1)
