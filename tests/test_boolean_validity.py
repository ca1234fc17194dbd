"""Boolean logic and validity predicates (the kernels a filter Expression is
made of): the oracle against the reference binary and the reference's own known-answer vectors (CPU), and
the C-ABI path against the oracle (GPU).
Reference tests followed: kernels/scalar_boolean_test.cc (And/Or/Xor/AndNot + Kleene truth tables),
kernels/scalar_validity_test.cc (IsValid / IsNull / nan_is_null / TrueUnlessNull / IsNan)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import arrow_oracle as ora

from .util import SEED

BINARY_OPS = ["and", "or", "xor", "and_not", "and_kleene", "or_kleene", "and_not_kleene"]

# scalar_boolean_test.cc: the 3 x 3 truth table every op is checked on
LEFT = pa.array([True, True, True, False, False, False, None, None, None])
RIGHT = pa.array([True, False, None, True, False, None, True, False, None])
KAT = {
    "and": [True, False, None, False, False, None, None, None, None],
    "and_kleene": [True, False, None, False, False, False, None, False, None],
    "or": [True, True, None, True, False, None, None, None, None],
    "or_kleene": [True, True, True, True, False, None, True, None, None],
    "xor": [False, True, None, True, False, None, None, None, None],
    "and_not": [False, True, None, False, False, None, None, None, None],
    "and_not_kleene": [False, True, None, False, False, False, False, None, None],
}


def _bools(n, null_p, seed, offset):
    rng = np.random.default_rng(seed)
    m = n + offset
    return pa.array(rng.random(m) < 0.5, mask=(rng.random(m) < null_p) if null_p else None).slice(offset)


def _floats(n, seed, offset, t=pa.float64()):
    rng = np.random.default_rng(seed)
    m = n + offset
    v = np.where(rng.random(m) < 0.15, np.nan, rng.uniform(-5, 5, m)).astype(t.to_pandas_dtype())
    return pa.array(v, type=t, mask=rng.random(m) < 0.2).slice(offset)


@pytest.mark.parametrize("op", BINARY_OPS)
def test_oracle_truth_tables(op):
    assert ora.boolean(op, LEFT, RIGHT).equals(pa.array(KAT[op], pa.bool_()))
    assert pc.call_function(op, [LEFT, RIGHT]).equals(pa.array(KAT[op], pa.bool_()))  # the KAT is the reference's


@pytest.mark.parametrize("op", BINARY_OPS)
@pytest.mark.parametrize("offset", [0, 3, 64])
def test_oracle_boolean_vs_reference_binary(op, offset):
    a, b = _bools(3000, 0.2, SEED, offset), _bools(3000, 0.3, SEED + 1, offset)
    assert ora.boolean(op, a, b).equals(pc.call_function(op, [a, b]))
    for s in (True, False, pa.scalar(None, pa.bool_())):
        assert ora.boolean(op, a, s).equals(pc.call_function(op, [a, s])), (op, s)
        assert ora.boolean(op, s, a).equals(pc.call_function(op, [s, a])), (op, s)
    assert ora.boolean("invert", a).equals(pc.invert(a))


@pytest.mark.parametrize("offset", [0, 5])
def test_oracle_validity_vs_reference_binary(offset):
    f = _floats(2000, SEED, offset)
    for op in ("is_valid", "is_null", "true_unless_null", "is_nan"):
        assert ora.validity_op(op, f).equals(pc.call_function(op, [f])), op
    assert ora.validity_op("is_null", f, True).equals(pc.is_null(f, nan_is_null=True))
    i = pa.array(np.arange(100), mask=np.arange(100) % 7 == 0).slice(offset)
    assert ora.validity_op("is_null", i, True).equals(pc.is_null(i, nan_is_null=True))
    assert ora.validity_op("is_valid", pa.array([1, 2, 3])).equals(pa.array([True] * 3))


def test_temporal_types_are_refused_not_relabelled():
    """timestamp / date / duration columns share int storage with the numeric kernels but
    must not run through them (casts would skip the unit rescale, compares would ignore units)."""
    import arrow_b200.compute as bc
    from arrow_b200.device import DeviceArray
    ts_s = DeviceArray(None, pa.timestamp("s"), 4, 0, 0, [None, None])
    ts_ms = DeviceArray(None, pa.timestamp("ms"), 4, 0, 0, [None, None])
    i64 = DeviceArray(None, pa.int64(), 4, 0, 0, [None, None])
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.cast(ts_s, pa.timestamp("ms"))
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.cast(DeviceArray(None, pa.date32(), 4, 0, 0, [None, None]), pa.date64())
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.cast(i64, pa.duration("s"))
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.equal(ts_s, ts_ms)
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.add(ts_s, i64)
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.sum(DeviceArray(None, pa.duration("s"), 4, 0, 0, [None, None]))
    assert bc.cast(ts_s, pa.timestamp("s")) is ts_s  # identical type: a no-op like the reference


# ------------------------------------------------------------------------------------------------
# GPU: the C-ABI kernels against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("op", BINARY_OPS)
def test_gpu_boolean_truth_tables(ctx, op):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    got = bc.call_function(op, [DeviceArray.from_arrow(LEFT, ctx), DeviceArray.from_arrow(RIGHT, ctx)]).to_arrow()
    assert got.equals(pa.array(KAT[op], pa.bool_()))


@pytest.mark.gpu
@pytest.mark.parametrize("op", BINARY_OPS)
@pytest.mark.parametrize("n,offset", [(1, 0), (63, 1), (4097, 7), (200003, 64), (0, 0)])
def test_gpu_boolean_vs_oracle(ctx, op, n, offset):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    a, b = _bools(n, 0.2, SEED + n, offset), _bools(n, 0.3, SEED + n + 1, offset + 2)
    da, db = DeviceArray.from_arrow(a, ctx), DeviceArray.from_arrow(b, ctx)
    assert bc.call_function(op, [da, db]).to_arrow().equals(ora.boolean(op, a, b))
    nn = _bools(n, 0.0, SEED + 5, offset)
    assert bc.call_function(op, [DeviceArray.from_arrow(nn, ctx), db]).to_arrow().equals(ora.boolean(op, nn, b))
    for s in (True, False, pa.scalar(None, pa.bool_())):
        assert bc.call_function(op, [da, s]).to_arrow().equals(ora.boolean(op, a, s)), (op, s)
        assert bc.call_function(op, [s, da]).to_arrow().equals(ora.boolean(op, s, a)), (op, s)
    assert bc.invert(da).to_arrow().equals(ora.boolean("invert", a))


@pytest.mark.gpu
@pytest.mark.parametrize("t", [pa.float64(), pa.float32()])
@pytest.mark.parametrize("n,offset", [(1, 0), (31, 3), (100003, 9), (0, 0)])
def test_gpu_validity_vs_oracle(ctx, t, n, offset):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    f = _floats(n, SEED + n, offset, t)
    df = DeviceArray.from_arrow(f, ctx)
    for op in ("is_valid", "is_null", "true_unless_null", "is_nan"):
        assert bc.call_function(op, [df]).to_arrow().equals(ora.validity_op(op, f)), op
    assert bc.is_null(df, nan_is_null=True).to_arrow().equals(ora.validity_op("is_null", f, True))
    i = pa.array(np.arange(n + offset, dtype=np.int32), mask=np.arange(n + offset) % 5 == 0).slice(offset)
    di = DeviceArray.from_arrow(i, ctx)
    assert bc.is_null(di).to_arrow().equals(ora.validity_op("is_null", i))
    assert bc.is_valid(di).to_arrow().equals(ora.validity_op("is_valid", i))
    with pytest.raises(pa.ArrowNotImplementedError):
        bc.is_nan(di)


@pytest.mark.gpu
def test_gpu_bincount(ctx):
    import ctypes as C
    from arrow_b200 import DeviceArray
    from arrow_b200.device import check
    rng = np.random.default_rng(SEED)
    for n, bins in ((0, 3), (1, 1), (100003, 9), (1 << 20, 4096)):
        ids = pa.array(rng.integers(0, bins, n, dtype=np.uint32))
        d = DeviceArray.from_arrow(ids, ctx)
        counts = (C.c_int64 * bins)()
        cd = d._c()
        check(ctx.lib.b2_bincount(ctx.handle, C.byref(cd), bins, counts, ctx.stream))
        assert list(counts) == np.bincount(ids.to_numpy(), minlength=bins).tolist()
    d = DeviceArray.from_arrow(pa.array(np.array([0, 5], dtype=np.uint32)), ctx)
    counts = (C.c_int64 * 3)()
    cd = d._c()
    with pytest.raises(pa.ArrowIndexError):
        check(ctx.lib.b2_bincount(ctx.handle, C.byref(cd), 3, counts, ctx.stream))
