"""if_else: the oracle against the reference binary (CPU), and the CUDA path against both (GPU).  Cases follow
TestIfElseKernel in kernels/scalar_if_else_test.cc:100-330: array/scalar shapes of all three arguments, nulls in each,
numeric promotion of left / right, boolean branches, sliced (offset) inputs, length mismatch."""
import itertools

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import arrow_oracle as ora
from tests.util import SEED, assert_equal, random_array

TYPES = [pa.int8(), pa.uint16(), pa.int32(), pa.int64(), pa.uint64(), pa.float32(), pa.float64(), pa.bool_()]


def operands(t, n, offset):
    cond = random_array(pa.bool_(), n, 0.1, SEED, offset=offset)
    if pa.types.is_boolean(t):
        left, right = random_array(t, n, 0.1, SEED + 1, offset=offset), random_array(t, n, 0.2, SEED + 2, offset=offset)
        scalars = (pa.scalar(True), pa.scalar(None, pa.bool_()))
    else:
        left = random_array(t, n, 0.1, SEED + 1, lo=0, hi=100, offset=offset)
        right = random_array(t, n, 0.2, SEED + 2, lo=0, hi=100, offset=offset)
        scalars = (pa.scalar(7, t), pa.scalar(None, t))
    return cond, left, right, scalars


def shapes(t, n, offset):
    cond, left, right, (s_valid, s_null) = operands(t, n, offset)
    for c in (cond, pa.scalar(True), pa.scalar(False), pa.scalar(None, pa.bool_())):
        for l, r in itertools.product((left, s_valid, s_null), (right, s_valid, s_null)):
            if any(isinstance(x, pa.Array) for x in (c, l, r)):
                yield c, l, r


@pytest.mark.parametrize("t", TYPES, ids=str)
def test_oracle_matches_the_reference_binary(t):
    for n, offset in ((0, 0), (1, 0), (67, 3), (1000, 0)):
        for c, l, r in shapes(t, n, offset):
            assert_equal(ora.if_else(c, l, r), pc.if_else(c, l, r), f"{t} n={n}")
    if not pa.types.is_boolean(t):   # promotion to the common numeric type
        cond, left, right, _ = operands(t, 200, 1)
        other = random_array(pa.int16(), 200, 0.1, SEED + 5, lo=-50, hi=50, offset=1)
        assert_equal(ora.if_else(cond, left, other), pc.if_else(cond, left, other))


@pytest.mark.gpu
@pytest.mark.parametrize("t", TYPES, ids=str)
def test_gpu_if_else_shapes(ctx, t):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray

    def dev(x):
        return DeviceArray.from_arrow(x, ctx) if isinstance(x, pa.Array) else x

    for n, offset in ((0, 0), (1, 0), (67, 3), (5000, 0), (5003, 5)):
        for c, l, r in shapes(t, n, offset):
            got = bc.if_else(dev(c), dev(l), dev(r)).to_arrow()
            assert_equal(got, ora.if_else(c, l, r), f"{t} n={n} offset={offset}")
            assert_equal(got, pc.if_else(c, l, r))


@pytest.mark.gpu
def test_gpu_if_else_large_promotion_and_errors(ctx):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    n = 3_000_001
    cond = random_array(pa.bool_(), n, 0.05, SEED)
    left = random_array(pa.int32(), n, 0.1, SEED + 1)
    right = random_array(pa.float64(), n, 0.1, SEED + 2)
    dc, dl, dr = (DeviceArray.from_arrow(x, ctx) for x in (cond, left, right))
    got = bc.if_else(dc, dl, dr).to_arrow()
    assert got.type == pa.float64()
    assert_equal(got, pc.if_else(cond, left, right))
    assert_equal(bc.if_else(dc, dl, pa.scalar(3, pa.int32())).to_arrow(), pc.if_else(cond, left, pa.scalar(3, pa.int32())))
    assert_equal(bc.if_else(dc, dl, 3).to_arrow(), pc.if_else(cond, left, 3))   # a Python int is an int64 scalar: promotes
    assert_equal(bc.call_function("if_else", [dc, dr, dr]).to_arrow(), pc.if_else(cond, right, right))
    with pytest.raises(pa.ArrowInvalid, match="same length"):
        bc.if_else(dc, dl.slice(0, 10), dl)
    with pytest.raises((pa.ArrowNotImplementedError, pa.ArrowTypeError)):
        bc.if_else(dl, dl, dl)
