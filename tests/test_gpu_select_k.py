"""select_k_unstable on one numeric column (csrc/select_k.cu; ArraySelector, kernels/vector_select_k.cc:157-232):
the result must be the first k rows of the stable sort (one of the permitted answers of the unstable selection), for
both the threshold path (k << n: sampled threshold, one compare pass, sort of the candidates) and the full-sort path.
The reference binary agrees on the selected VALUES whenever k does not reach into the NaNs / nulls."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from tests.util import SEED, random_array

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t", [pa.int64(), pa.int32(), pa.uint16(), pa.float64(), pa.float32()], ids=str)
@pytest.mark.parametrize("order", ["ascending", "descending"])
def test_select_k_equals_sort_prefix(ctx, t, order):
    for n, null_p, hi in ((3_000_000, 0.1, 10**6), (2_000_000, 0.0, 50), (5000, 0.2, 100)):
        arr = random_array(t, n, null_p, SEED + n, lo=0 if pa.types.is_unsigned_integer(t) else -hi, hi=hi, offset=3)
        if pa.types.is_floating(t):  # a few NaNs: they sort after the values, before the nulls
            v = arr.to_numpy(zero_copy_only=False).copy()
            v[::1001] = np.nan
            arr = pa.array(v, t, mask=np.asarray(arr.is_null()))
        d = DeviceArray.from_arrow(arr, ctx)
        full = bc.array_sort_indices(d, order).to_arrow()
        for k in (0, 1, 10, 4097, n // 10, n + 5):
            got = bc.select_k_unstable(d, k, [("x", order)]).to_arrow()
            assert got.equals(full.slice(0, min(k, n))), f"{t} {order} n={n} k={k}"
        # against the reference binary: same selected values while k stays inside the plain values
        k = 1000
        want = pc.select_k_unstable(arr, k, [("x", order)])
        got = bc.call_function("select_k_unstable", [d], pc.SelectKOptions(k, [("x", order)])).to_arrow()
        assert pc.take(arr, got).equals(pc.take(arr, want))
    with pytest.raises(pa.ArrowInvalid, match="nonnegative"):
        bc.select_k_unstable(d, -1)


def test_select_k_nulls_first_and_skew(ctx):
    n = 2_500_000
    rng = np.random.default_rng(SEED)
    # heavy skew: 99 % of the rows hold one value, so the sampled threshold admits almost everything -> full sort fallback
    v = np.where(rng.random(n) < 0.99, 7, rng.integers(-1000, 1000, n))
    arr = pa.array(v, pa.int64(), mask=rng.random(n) < 0.05)
    d = DeviceArray.from_arrow(arr, ctx)
    for order in ("ascending", "descending"):
        for placement in ("at_end", "at_start"):
            full = bc.array_sort_indices(d, order, placement).to_arrow()
            for k in (5, 30000):
                got = bc.select_k_unstable(d, k, [("x", order)], null_placement=placement).to_arrow()
                assert got.equals(full.slice(0, k)), f"{order} {placement} k={k}"
