"""Sender-side kernels of the multi-GPU exchange: b2_range_split (stable split of (value, row) pairs by range id)."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from tests.util import SEED, random_array

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t", [pa.int64(), pa.float64(), pa.int16(), pa.uint32()], ids=str)
@pytest.mark.parametrize("n,n_split,null_p,offset", [(0, 1, 0.0, 0), (1, 0, 0.0, 0), (2049, 1, 0.1, 3), (100_003, 3, 0.1, 1), (300_000, 7, 0.0, 0),
                                                     (50_000, 15, 0.5, 5)])
def test_range_split_is_a_stable_partition(ctx, t, n, n_split, null_p, offset):
    from arrow_b200 import DeviceArray, _cabi as cabi
    from arrow_b200.device import check
    from oracle import arrow_oracle as ora
    vals = random_array(t, n, null_p, SEED + n, lo=0, hi=1000, offset=offset)
    rng = np.random.default_rng(SEED)
    sp = np.sort(rng.integers(0, 1000, n_split)).astype(t.to_pandas_dtype())
    splitters = pa.array(sp, t)
    row_base = 12345
    dv, ds = DeviceArray.from_arrow(vals, ctx), DeviceArray.from_arrow(splitters, ctx)
    counts = (C.c_int64 * (n_split + 2))()
    cv, cs, ov, orows = dv._c(), ds._c(), cabi.B2Array(), cabi.B2Array()
    check(ctx.lib.b2_range_split(ctx.handle, C.byref(cv), C.byref(cs), 0, row_base, C.byref(ov), C.byref(orows), counts, ctx.stream))
    got_v = DeviceArray._from_c(ctx, ov, t).to_arrow()
    got_r = DeviceArray._from_c(ctx, orows, pa.uint32()).to_arrow().to_numpy().astype(np.int64)
    v, valid = ora.values(vals), ora.validity(vals)
    ids = np.where(valid, np.searchsorted(sp, v, side="right"), n_split + 1)
    order = np.argsort(ids, kind="stable")
    assert list(counts) == np.bincount(ids, minlength=n_split + 2).tolist()
    assert (got_r == order + row_base).all()
    n_valid = int(valid.sum())
    got_vals = got_v.to_numpy(zero_copy_only=False)
    if pa.types.is_floating(t):
        assert np.array_equal(got_vals[:n_valid], v[order][:n_valid], equal_nan=True)
    else:
        assert (got_vals[:n_valid] == v[order][:n_valid]).all()
