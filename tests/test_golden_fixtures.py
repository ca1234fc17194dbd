"""Golden vectors produced by the reference binary (tests/golden/make_reference_fixtures.py ->
tests/golden/reference_fixtures.arrow): the oracle (CPU) and the CUDA path through the C-ABI (GPU)
must reproduce every stored result without calling the reference at test time."""
import json
import os

import numpy as np
import pyarrow as pa
import pytest

from oracle import arrow_oracle as ora
from tests.util import assert_equal

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    with pa.OSFile(os.path.join(HERE, "golden", "reference_fixtures.arrow"), "rb") as f:
        batch = pa.ipc.open_stream(f).read_next_batch()
    cols = {n: batch.column(i).values for i, n in enumerate(batch.schema.names)}
    cases = json.loads(batch.schema.metadata[b"cases"])
    return [(c, [cols[n] for n in c["inputs"]], cols[c["output"]]) for c in cases]


CASES = load()
FLOAT_TOL = 1e-12  # sums / means: the reference adds pairwise; relative to the sum of magnitudes


def evaluate(api, case, ins, to_host, dev):
    """`api` is either the oracle adapter or arrow_b200.compute; `dev` moves an input to where `api` wants it."""
    fn, o = case["function"], case["options"]
    a = [dev(x) for x in ins]
    if fn == "filter":
        return to_host(api.filter(a[0], a[1], o["null_selection_behavior"]))
    if fn == "take":
        return to_host(api.take(a[0], a[1]))
    if fn == "cast":
        return to_host(api.cast(a[0], pa.type_for_alias(o["to"]), safe=o["safe"]))
    if fn in ("add", "subtract", "multiply"):
        return to_host(api.arithmetic(fn, a[0], a[1]))
    if fn in ("equal", "less", "greater_equal"):
        return to_host(api.compare(fn, a[0], a[1]))
    if fn == "array_sort_indices":
        return to_host(api.sort_indices(a[0], o["order"], o["null_placement"]))
    if fn == "unique":
        return to_host(api.unique(a[0]))
    if fn.startswith("value_counts."):
        v, c = api.value_counts(a[0])
        return to_host(v if fn.endswith("values") else c)
    if fn.startswith("dictionary_encode."):
        d = api.dictionary_encode(a[0], o["null_encoding"])
        return d.indices if fn.endswith("indices") else d.dictionary
    if fn in ("sum", "mean"):
        s = getattr(api, "scalar_" + fn)(a[0], o["skip_nulls"], o["min_count"])
        return pa.array([s.as_py()], s.type)
    if fn == "min_max":
        s = api.scalar_min_max(a[0], o["skip_nulls"], o["min_count"])
        return pa.array([s["min"].as_py(), s["max"].as_py()], ins[0].type)
    if fn.startswith("group_by."):
        col = fn.split(".", 1)[1]
        uniq, outs = api.group_by([a[0]], [("hash_" + f, a[1], None) for f in ("sum", "count", "min", "max")])
        t = pa.table({"k": to_host(uniq[0]), **{"v_" + f: to_host(x) for f, x in zip(("sum", "count", "min", "max"), outs)}}).sort_by("k")
        return t[col].combine_chunks()
    raise AssertionError(fn)


def check(case, got, want, ins):
    if case["function"] in ("sum", "mean") and pa.types.is_floating(ins[0].type) and want[0].is_valid:
        scale = float(np.abs(ora.values(ins[0])[ora.validity(ins[0])].astype(np.float64)).sum())
        if case["function"] == "mean":
            scale /= max(1, len(ins[0]) - ins[0].null_count)
        assert got[0].is_valid and abs(got[0].as_py() - want[0].as_py()) <= FLOAT_TOL * (scale or 1.0), case["name"]
    else:
        assert_equal(got, want, case["name"])  # bit-exact (Array::Equals semantics: bytes under nulls are free)


class OracleApi:
    filter, take, unique = staticmethod(ora.filter), staticmethod(ora.take), staticmethod(ora.unique)
    scalar_sum, scalar_mean, scalar_min_max = staticmethod(ora.scalar_sum), staticmethod(ora.scalar_mean), staticmethod(ora.scalar_min_max)
    arithmetic, compare, sort_indices, group_by = (staticmethod(ora.arithmetic), staticmethod(ora.compare),
                                                    staticmethod(ora.sort_indices), staticmethod(ora.group_by))
    dictionary_encode = staticmethod(ora.dictionary_encode)

    @staticmethod
    def cast(a, to, safe):
        return ora.cast_array(a, to, safe)

    @staticmethod
    def value_counts(a):
        s = ora.value_counts(a)
        return s.field("values"), s.field("counts")


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c[0]["name"] for c in CASES])
def test_oracle_reproduces_reference_fixture(i):
    case, ins, want = CASES[i]
    check(case, evaluate(OracleApi, case, ins, lambda x: x, lambda x: x), want, ins)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)), ids=[c[0]["name"] for c in CASES])
def test_cuda_path_reproduces_reference_fixture(ctx, i):
    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray

    class DeviceApi:
        filter, take, cast, unique, value_counts, group_by = bc.filter, bc.take, bc.cast, bc.unique, bc.value_counts, bc.group_by
        scalar_sum, scalar_mean, scalar_min_max = bc.sum, bc.mean, bc.min_max

        @staticmethod
        def arithmetic(op, l, r):
            return getattr(bc, op)(l, r)

        compare = arithmetic

        @staticmethod
        def sort_indices(a, order, null_placement):
            return bc.array_sort_indices(a, order, null_placement)

        @staticmethod
        def dictionary_encode(a, mode):
            return bc.dictionary_encode(a, mode).to_arrow()

    case, ins, want = CASES[i]
    got = evaluate(DeviceApi, case, ins, lambda x: x.to_arrow(), lambda x: DeviceArray.from_arrow(x, ctx))
    check(case, got, want, ins)
