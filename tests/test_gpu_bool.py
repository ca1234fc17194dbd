"""Boolean (bit-packed) value columns through Filter and Take: the kIsBoolean paths of
PrimitiveFilterImpl / Gather (vector_selection_filter_internal.cc:478-480, gather_internal.h:172-251)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from oracle import arrow_oracle as ora
from tests.util import SEED, assert_equal, random_array

pytestmark = pytest.mark.gpu


def dev(a, ctx):
    return DeviceArray.from_arrow(a, ctx)


def test_bool_filter(ctx):
    for n, off in ((5, 0), (1024, 3), (70001, 5), (4096 * 3, 0)):
        for null_p in (0.0, 0.1, 1.0):
            vals = random_array(pa.bool_(), n, null_p, SEED + n, hi=0.4, offset=off)
            for true_p, mask_null in ((0.5, 0.0), (0.02, 0.05), (1.0, 0.3), (0.0, 0.0)):
                mask = random_array(pa.bool_(), n, mask_null, SEED + 7, hi=true_p, offset=(off * 3) % 7)
                for ns in ("drop", "emit_null"):
                    got = bc.filter(dev(vals, ctx), dev(mask, ctx), ns)
                    want = pc.filter(vals, mask, null_selection_behavior=ns)
                    assert_equal(got.to_arrow(), want, f"n={n} {null_p} {true_p} {ns}")
                    assert_equal(got.to_arrow(), ora.filter(vals, mask, ns))
                    assert got.null_count == want.null_count


def test_bool_take(ctx):
    for null_p in (0.0, 0.1, 0.9):
        vals = random_array(pa.bool_(), 3001, null_p, SEED, hi=0.5, offset=5)
        for it in (pa.int8(), pa.uint16(), pa.int32(), pa.int64(), pa.uint64()):
            hi = min(3000, np.iinfo(it.to_pandas_dtype()).max)
            for n_idx in (17, 9000):
                idx = random_array(it, n_idx, null_p, SEED + 3, lo=0, hi=hi, offset=2)
                got = bc.take(dev(vals, ctx), dev(idx, ctx))
                want = pc.take(vals, idx)
                assert_equal(got.to_arrow(), want, f"{it} {null_p} {n_idx}")
                assert_equal(got.to_arrow(), ora.take(vals, idx))
    with pytest.raises(pa.ArrowIndexError, match="out of bounds"):
        bc.take(dev(pa.array([True, False]), ctx), dev(pa.array([0, 2]), ctx))
