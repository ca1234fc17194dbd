"""Multi-key SortIndices (record batch / table sort): the oracle against the reference binary (CPU) and the CUDA
composition -- stable single-key radix sorts from the last key to the first, the permutation riding along as the
payload -- against both (GPU).  Cases follow TestRecordBatchSortIndices / TestTableSortIndices in
kernels/vector_sort_test.cc:1105-1450: mixed orders, nulls and NaNs in several keys, duplicates, null placement."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import arrow_oracle as ora
from tests.util import SEED, random_array


def table(n, seed=SEED):
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 4, n).astype(np.float64)
    f[rng.random(n) < 0.15] = np.nan
    return {
        "a": random_array(pa.int32(), n, 0.1, seed, lo=0, hi=5),
        "b": pa.array(f, pa.float64(), mask=rng.random(n) < 0.1),
        "c": random_array(pa.int64(), n, 0.0, seed + 2, lo=-3, hi=3),
        "d": random_array(pa.uint8(), n, 0.2, seed + 3, lo=0, hi=2),
    }


KEYS = [
    [("a", "ascending"), ("b", "descending")],
    [("b", "ascending"), ("a", "ascending"), ("c", "descending")],
    [("d", "descending"), ("c", "ascending"), ("b", "descending"), ("a", "descending")],
    [("c", "ascending")],
]


def reference(cols, keys, placement):
    return pc.sort_indices(pa.table(cols), sort_keys=keys, null_placement=placement)


@pytest.mark.parametrize("placement", ["at_end", "at_start"])
def test_oracle_matches_the_reference_binary(placement):
    for n in (0, 1, 2, 500):
        cols = table(n)
        for keys in KEYS:
            got = ora.sort_indices_multi(cols, keys, placement)
            assert got.equals(reference(cols, keys, placement)), (n, keys)


@pytest.mark.gpu
@pytest.mark.parametrize("placement", ["at_end", "at_start"])
def test_gpu_multi_key_sort(ctx, placement):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    for n in (0, 1, 37, 20_000):
        cols = table(n)
        dcols = {k: DeviceArray.from_arrow(v, ctx) for k, v in cols.items()}
        for keys in KEYS:
            got = bc.sort_indices(dcols, sort_keys=keys, null_placement=placement).to_arrow()
            assert got.equals(ora.sort_indices_multi(cols, keys, placement)), (n, keys)
            assert got.equals(reference(cols, keys, placement)), (n, keys)


@pytest.mark.gpu
def test_gpu_multi_key_sort_large_and_errors(ctx):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray
    n = 2_000_003
    cols = table(n, SEED + 9)
    sliced = {k: v.slice(3) for k, v in cols.items()}          # non-zero offsets on every key
    dcols = {k: DeviceArray.from_arrow(v, ctx) for k, v in sliced.items()}
    keys = [("a", "descending"), ("b", "ascending"), ("c", "ascending")]
    got = bc.sort_indices(dcols, sort_keys=keys).to_arrow()
    assert got.equals(reference(sliced, keys, "at_end"))
    with pytest.raises(pa.ArrowInvalid, match="sort keys"):
        bc.sort_indices(dcols, sort_keys=[])
    with pytest.raises(pa.ArrowInvalid, match="No match"):
        bc.sort_indices(dcols, sort_keys=[("zz", "ascending")])
