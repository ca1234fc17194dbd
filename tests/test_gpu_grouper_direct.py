"""Grouper, direct-addressed mode (grouper.cu): one integer key column whose values span <= 2^24 is grouped through
L2-resident {id, first_row} arrays instead of the hash table.  Same contract as the hash mode -- TestGrouper::ValidateConsume
(row/grouper_test.cc:736-760): ids in first-occurrence order, uniques prefix-stable, Lookup never inserts -- checked
against the oracle and against the hash mode (B2_GROUPER_DIRECT=0) on the same batches."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray
from oracle import arrow_oracle as ora
from tests.util import SEED, assert_equal, random_array

pytestmark = pytest.mark.gpu


def dev(arr, ctx):
    return DeviceArray.from_arrow(arr, ctx)


def first_occurrence_ids(keys: np.ndarray):
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    return rank[inv], keys[np.sort(first)]


def run_batches(kt, batches, ctx, lookups=()):
    g, og = bc.Grouper([kt], ctx), ora.Grouper([kt])
    prev = None
    for b, keys in enumerate(batches):
        ids = g.consume(dev(keys, ctx)).to_arrow()
        assert_equal(ids, og.consume(keys), f"{kt} batch {b}")
        uniq = g.get_uniques()[0].to_arrow()
        assert g.num_groups == og.num_groups == len(uniq)
        assert pc.take(uniq, ids).equals(keys)
        if prev is not None:
            assert uniq.slice(0, len(prev)).equals(prev)
        prev = uniq
    for look in lookups:
        assert_equal(g.lookup(dev(look, ctx)).to_arrow(), og.lookup(look))
    return g.get_uniques()[0].to_arrow()


@pytest.mark.parametrize("kt", [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint16(), pa.uint64()], ids=str)
def test_signed_windows_nulls_and_growth(ctx, kt, monkeypatch):
    signed = pa.types.is_signed_integer(kt)
    lo0, hi0 = (-50, 60) if signed else (1000 if kt.bit_width > 8 else 10, 1100 if kt.bit_width > 8 else 100)
    step = 20 if kt.bit_width == 8 else 3000
    batches = [
        random_array(kt, 4000, 0.05, SEED, lo=lo0, hi=hi0),
        random_array(kt, 4000, 0.0, SEED + 1, lo=lo0, hi=hi0 + step),             # window grows upward
        random_array(kt, 4000, 0.3, SEED + 2, lo=lo0 - (step if signed else 5), hi=hi0),  # ... and downward
        pa.array([None] * 37, kt),                                               # all-null batch
        random_array(kt, 100, 0.0, SEED + 3, lo=lo0, hi=hi0),                    # nothing new
    ]
    looks = [random_array(kt, 700, 0.1, SEED + 9, lo=lo0 - (100 if kt.bit_width > 8 and signed else 0), hi=hi0 + 2 * step)]
    direct = run_batches(kt, batches, ctx, looks)
    monkeypatch.setenv("B2_GROUPER_DIRECT", "0")
    assert run_batches(kt, batches, ctx, looks).equals(direct)


def test_leaves_direct_mode_when_a_batch_is_too_wide(ctx):
    kt = pa.int64()
    rng = np.random.default_rng(SEED)
    narrow = pa.array(rng.integers(0, 5000, 20000), kt)
    wide = pa.array(np.concatenate([rng.integers(0, 5000, 5000), rng.integers(-2**62, 2**62, 5000)]), kt)
    after = pa.array(rng.integers(-10, 6000, 20000), kt)
    look = pa.array(np.concatenate([rng.integers(-100, 7000, 300), wide.to_numpy()[-50:]]), kt)
    run_batches(kt, [narrow, wide, after], ctx, [look])
    # the extreme ends of the domain in one batch: the window arithmetic must not wrap
    ends = pa.array([np.iinfo(np.int64).min, np.iinfo(np.int64).max, 0, None, np.iinfo(np.int64).max], kt)
    run_batches(kt, [ends, narrow], ctx, [look])
    run_batches(pa.uint64(), [pa.array([2**64 - 1, 2**64 - 3, 2**64 - 1, None], pa.uint64()),
                              pa.array([0, 5, 2**64 - 2], pa.uint64())], ctx)


def test_first_batch_sparse_goes_to_the_hash_table(ctx):
    # 100 rows spanning 10M values: not worth a direct table, and the later dense batches keep working
    kt = pa.int32()
    rng = np.random.default_rng(SEED + 4)
    run_batches(kt, [pa.array(rng.integers(0, 10_000_000, 100), kt), pa.array(rng.integers(0, 3000, 50000), kt)], ctx)


@pytest.mark.parametrize("n,groups", [(3_000_000, 1_000_000), (4_000_000, 17)])
def test_large_batches_first_occurrence_order(ctx, n, groups):
    rng = np.random.default_rng(SEED + n)
    k1, k2 = rng.integers(-groups // 2, groups // 2, n), rng.integers(-groups, groups, n // 2)
    g = bc.Grouper([pa.int64()], ctx)
    ids1 = g.consume(dev(pa.array(k1, pa.int64()), ctx)).to_arrow().to_numpy()
    ids2 = g.consume(dev(pa.array(k2, pa.int64()), ctx)).to_arrow().to_numpy()
    want, uniq = first_occurrence_ids(np.concatenate([k1, k2]))
    assert np.array_equal(ids1, want[:n]) and np.array_equal(ids2, want[n:])
    assert np.array_equal(g.get_uniques()[0].to_arrow().to_numpy(), uniq)


def test_vector_hash_kernels_ride_the_direct_mode(ctx):
    v = random_array(pa.int32(), 200_000, 0.02, SEED, lo=-300, hi=900)
    dv = dev(v, ctx)
    assert_equal(bc.unique(dv).to_arrow(), pc.unique(v))
    enc = bc.dictionary_encode(dv).to_arrow()
    assert enc.equals(pc.dictionary_encode(v))
    assert bc.value_counts_to_struct(*bc.value_counts(dv)).equals(pc.value_counts(v))
