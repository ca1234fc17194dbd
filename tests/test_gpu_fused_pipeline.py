"""b2_take_cast_arith: the fused form of add(cast(take(values, indices), T, safe=False), other) (BASELINE configs[1])
must equal the three reference kernels applied in sequence -- values, validity, null count and the IndexError."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from oracle import arrow_oracle as ora

from .util import SEED, random_array

pytestmark = pytest.mark.gpu


def _expect(values, idx, to, op, other):
    return ora.arithmetic(op, ora.cast_array(ora.take(values, idx), to, safe=False), other)


@pytest.mark.parametrize("vt", [pa.float64(), pa.float32(), pa.int64(), pa.int32()], ids=str)
@pytest.mark.parametrize("it", [pa.int64(), pa.int32(), pa.uint32(), pa.uint64()], ids=str)
@pytest.mark.parametrize("to", [pa.float32(), pa.float64()], ids=str)
def test_fused_equals_three_kernels(ctx, vt, it, to):
    for n_v, n, vnull, inull, onull, off in ((1000, 1, 0.1, 0.0, 0.1, 0), (5000, 4099, 0.1, 0.05, 0.1, 3), (70000, 200003, 0.0, 0.0, 0.0, 1),
                                             (70000, 131072, 0.3, 0.0, 0.0, 0)):
        values = random_array(vt, n_v, vnull, SEED + n, lo=-1000, hi=1000, offset=off)
        idx = random_array(it, n, inull, SEED + 1, lo=0, hi=n_v - 1, offset=off)
        other = random_array(to, n, onull, SEED + 2, lo=-10, hi=10, offset=2 * off)
        dv, di, do = (DeviceArray.from_arrow(x, ctx) for x in (values, idx, other))
        for op in ("add", "subtract", "multiply"):
            got = bc.take_cast_arith(dv, di, to, op, do)
            want = _expect(values, idx, to, op, other)
            assert got.to_arrow().equals(want), f"{vt} {it} {to} {op} n={n}"
            assert got.null_count == want.null_count
            # and against the reference binary itself
            ref = pc.call_function(op, [pc.cast(pc.take(values, idx), to, safe=False), other])
            assert got.to_arrow().equals(ref)


def test_fused_bit_exact_and_errors(ctx):
    rng = np.random.default_rng(SEED)
    n = 100_003
    values = pa.array(rng.uniform(0, 1e6, n), mask=rng.random(n) < 0.1)
    idx = pa.array(rng.integers(0, n, n, dtype=np.int64))
    other = pa.array(rng.uniform(0, 1e6, n).astype(np.float32), mask=rng.random(n) < 0.1)
    dv, di, do = (DeviceArray.from_arrow(x, ctx) for x in (values, idx, other))
    got = bc.take_cast_arith(dv, di, pa.float32(), "add", do).to_arrow()
    three = bc.add(bc.cast(bc.take(dv, di), pa.float32(), safe=False), do).to_arrow()
    assert got.equals(three)
    g, t = got.fill_null(0).to_numpy(), three.fill_null(0).to_numpy()
    assert (g.view(np.uint32) == t.view(np.uint32)).all()   # same bits, not just equal values
    bad = pa.array([0, 5, n, 1], pa.int64())
    with pytest.raises(pa.ArrowIndexError) as want:
        pc.take(values, bad)
    with pytest.raises(pa.ArrowIndexError) as err:
        bc.take_cast_arith(dv, DeviceArray.from_arrow(bad, ctx), pa.float32(), "add", DeviceArray.from_arrow(other.slice(0, 4), ctx))
    assert str(err.value) == str(want.value)
    # operand shapes the fused kernel does not cover fall back to the three calls with the same result
    u8 = pa.array(rng.integers(0, 200, n, dtype=np.uint8))
    d8 = DeviceArray.from_arrow(u8, ctx)
    assert bc.take_cast_arith(d8, di, pa.float32(), "add", do).to_arrow().equals(_expect(u8, idx, pa.float32(), "add", other))
    assert len(bc.take_cast_arith(dv, DeviceArray.from_arrow(idx.slice(0, 0), ctx), pa.float32(), "add",
                                  DeviceArray.from_arrow(other.slice(0, 0), ctx))) == 0
