"""Shared test helpers: seeded random arrays in the shape of the reference's
RandomArrayGenerator (cpp/src/arrow/testing/random.cc) and KAT loading."""
import json
import os

import numpy as np
import pyarrow as pa

SEED = 0x0FF1CE  # kRandomSeed, compute/kernels/test_util_internal.h:135
HERE = os.path.dirname(os.path.abspath(__file__))

NUMERIC_TYPES = [pa.int8(), pa.uint8(), pa.int16(), pa.uint16(), pa.int32(), pa.uint32(), pa.int64(), pa.uint64(),
                 pa.float32(), pa.float64()]
INT_TYPES = NUMERIC_TYPES[:8]
TYPE_BY_NAME = {str(t): t for t in NUMERIC_TYPES}
TYPE_BY_NAME.update({"float32": pa.float32(), "float64": pa.float64()})


def kat():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)


def from_json(t, vals):
    """ArrayFromJSON for numeric/bool lists where NaN is spelled "NaN"."""
    vals = [float("nan") if v == "NaN" else v for v in vals]
    if pa.types.is_boolean(t):
        vals = [None if v is None else bool(v) for v in vals]
    return pa.array(vals, type=t)


def np_dtype(t):
    return t.to_pandas_dtype()


def random_array(t, n, null_probability=0.0, seed=SEED, lo=None, hi=None, offset=0):
    """Random array of type t; `offset` > 0 returns a slice of a longer array so kernels see a
    non-zero, non-byte-aligned ArraySpan.offset (the reference re-runs every test on slices)."""
    rng = np.random.default_rng(seed)
    m = n + offset
    if pa.types.is_boolean(t):
        vals = rng.random(m) < (0.5 if hi is None else hi)
        arr = pa.array(vals, type=t, mask=_mask(rng, m, null_probability))
    elif pa.types.is_floating(t):
        vals = rng.uniform(-1e6 if lo is None else lo, 1e6 if hi is None else hi, m).astype(np_dtype(t))
        arr = pa.array(vals, type=t, mask=_mask(rng, m, null_probability))
    elif pa.types.is_integer(t):
        info = np.iinfo(np_dtype(t))
        l = info.min if lo is None else max(lo, info.min)
        h = info.max if hi is None else min(hi, info.max)
        vals = rng.integers(l, h, m, dtype=np_dtype(t), endpoint=True)
        arr = pa.array(vals, type=t, mask=_mask(rng, m, null_probability))
    elif pa.types.is_string(t) or pa.types.is_large_string(t) or pa.types.is_binary(t) or pa.types.is_large_binary(t):
        lens = rng.integers(0 if lo is None else lo, 32 if hi is None else hi, m, endpoint=True)
        chars = rng.integers(97, 122, int(lens.sum()), dtype=np.uint8, endpoint=True)
        offs = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        ow = np.int64 if (pa.types.is_large_string(t) or pa.types.is_large_binary(t)) else np.int32
        mask = _mask(rng, m, null_probability)
        vbuf = None if mask is None else pa.py_buffer(np.packbits(~mask, bitorder="little").tobytes())
        arr = pa.Array.from_buffers(t, m, [vbuf, pa.py_buffer(offs.astype(ow).tobytes()), pa.py_buffer(chars.tobytes())],
                                    null_count=-1 if mask is not None else 0)
    else:
        raise NotImplementedError(str(t))
    return arr.slice(offset) if offset else arr


def _mask(rng, m, p):
    if p <= 0:
        return None
    return rng.random(m) < p


def assert_equal(got: pa.Array, want: pa.Array, msg=""):
    """Array::Equals: type, length, validity and the values AT VALID SLOTS."""
    assert got.type == want.type, f"{msg} type {got.type} != {want.type}"
    assert len(got) == len(want), f"{msg} length {len(got)} != {len(want)}"
    got.validate(full=True)
    if not got.equals(want):
        import itertools
        bad = [i for i, (a, b) in enumerate(itertools.islice(zip(got, want), 0, 100000)) if a != b and not (
            a.is_valid and b.is_valid and a.as_py() != a.as_py() and b.as_py() != b.as_py())][:5]
        raise AssertionError(f"{msg} arrays differ at {bad}: got {[got[i] for i in bad]} want {[want[i] for i in bad]}")


def equal_nan(got: pa.Array, want: pa.Array) -> bool:
    return got.equals(want) or (len(got) == len(want) and got.type == want.type and
                                got.is_valid().equals(want.is_valid()) and
                                all((a.as_py() == b.as_py()) or (a.as_py() != a.as_py() and b.as_py() != b.as_py())
                                    for a, b in zip(got, want) if a.is_valid))
