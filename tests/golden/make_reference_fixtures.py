#!/usr/bin/env python
"""Generates tests/golden/reference_fixtures.arrow: seeded inputs and the outputs the REFERENCE BINARY
(pyarrow 24.0.0 = libarrow_compute.so.2400 of apache/arrow) produces for them, one
record per operation.  The file is an Arrow IPC stream whose schema metadata lists the cases; every case
stores its input arrays and the expected result as columns of a one-row-group table (lists), so the
parity tests (tests/test_golden_fixtures.py) need neither the reference binary's compute module nor this
script at run time.  Re-run after changing the case list:  python tests/golden/make_reference_fixtures.py
"""
import json
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.util import SEED, random_array  # noqa: E402

N = 333  # not a multiple of 64: the last bitmap word is partial


def cases():
    out = []
    i64 = random_array(pa.int64(), N, 0.1, SEED, lo=-50, hi=50, offset=3)
    f64 = random_array(pa.float64(), N, 0.1, SEED + 1, offset=5)
    u16 = random_array(pa.uint16(), N, 0.0, SEED + 2, lo=0, hi=40)
    f32 = random_array(pa.float32(), N, 0.2, SEED + 3, lo=-1000, hi=1000)
    mask = random_array(pa.bool_(), N, 0.05, SEED + 4, hi=0.5, offset=2)
    idx = random_array(pa.int32(), 200, 0.05, SEED + 5, lo=0, hi=N - 1)
    strs = random_array(pa.string(), N, 0.1, SEED + 6, lo=0, hi=12)
    for name, v in (("int64", i64), ("float64", f64), ("string", strs)):
        for ns in ("drop", "emit_null"):
            out.append((f"filter/{name}/{ns}", "filter", {"null_selection_behavior": ns}, [v, mask], pc.filter(v, mask, null_selection_behavior=ns)))
        out.append((f"take/{name}", "take", {}, [v, idx], pc.take(v, idx)))
    for v, to in ((f64, pa.float32()), (i64, pa.int8()), (u16, pa.float64()), (f32, pa.int64())):
        out.append((f"cast/{v.type}->{to}", "cast", {"to": str(to), "safe": False}, [v], pc.cast(v, to, safe=False)))
    i64b = random_array(pa.int64(), N, 0.1, SEED + 7, lo=-50, hi=50)
    for op in ("add", "subtract", "multiply"):
        out.append((f"{op}/int64", op, {}, [i64, i64b], getattr(pc, op)(i64, i64b)))
    out.append(("add/float32+float64", "add", {}, [f32, f64], pc.add(f32, f64)))
    for op in ("equal", "less", "greater_equal"):
        out.append((f"{op}/int64", op, {}, [i64, i64b], getattr(pc, op)(i64, i64b)))
    for v in (i64, f64, u16):
        for order in ("ascending", "descending"):
            for np_ in ("at_end", "at_start"):
                out.append((f"sort/{v.type}/{order}/{np_}", "array_sort_indices", {"order": order, "null_placement": np_}, [v],
                            pc.array_sort_indices(v, order=order, null_placement=np_)))
    for v in (i64, u16):
        out.append((f"unique/{v.type}", "unique", {}, [v], pc.unique(v)))
        vc = pc.value_counts(v)
        out.append((f"value_counts/{v.type}/values", "value_counts.values", {}, [v], vc.field("values")))
        out.append((f"value_counts/{v.type}/counts", "value_counts.counts", {}, [v], vc.field("counts")))
        for mode in ("mask", "encode"):
            de = pc.dictionary_encode(v, null_encoding=mode)
            out.append((f"dictionary_encode/{v.type}/{mode}/indices", "dictionary_encode.indices", {"null_encoding": mode}, [v], de.indices))
            out.append((f"dictionary_encode/{v.type}/{mode}/dictionary", "dictionary_encode.dictionary", {"null_encoding": mode}, [v], de.dictionary))
    for v in (i64, u16, f64):
        for skip in (True, False):
            o = {"skip_nulls": skip, "min_count": 1}
            out.append((f"sum/{v.type}/skip={skip}", "sum", o, [v], pa.array([pc.sum(v, **o).as_py()], pc.sum(v).type)))
            out.append((f"mean/{v.type}/skip={skip}", "mean", o, [v], pa.array([pc.mean(v, **o).as_py()], pa.float64())))
            mm = pc.min_max(v, **o)
            out.append((f"min_max/{v.type}/skip={skip}", "min_max", o, [v], pa.array([mm["min"].as_py(), mm["max"].as_py()], v.type)))
    # group-by: keys with nulls, int64 + float values; result sorted by key (the reference's own test convention)
    keys = random_array(pa.int64(), N, 0.05, SEED + 8, lo=0, hi=20)
    t = pa.table({"k": keys, "v": i64}).group_by("k", use_threads=False).aggregate([("v", "sum"), ("v", "count"), ("v", "min"), ("v", "max")]).sort_by("k")
    for col in ("k", "v_sum", "v_count", "v_min", "v_max"):
        out.append((f"group_by/int64/{col}", "group_by." + col, {}, [keys, i64], t[col].combine_chunks()))
    return out


def main():
    rows = cases()
    fields, arrays, meta, pool = [], [], [], {}

    def column(a, name):
        a = a.combine_chunks() if isinstance(a, pa.ChunkedArray) else a
        # every array is stored as ONE list value so arrays of different lengths share a record batch
        fields.append(pa.field(name, pa.list_(a.type)))
        arrays.append(pa.ListArray.from_arrays(pa.array([0, len(a)], pa.int32()), pa.concat_arrays([a])))
        return name

    for j, (name, fn, opts, ins, want) in enumerate(rows):
        in_cols = []
        for a in ins:  # inputs are shared between cases: store each distinct array once
            if id(a) not in pool:
                pool[id(a)] = column(a, f"in{len(pool)}")
            in_cols.append(pool[id(a)])
        meta.append({"name": name, "function": fn, "options": opts, "inputs": in_cols, "output": column(want, f"out{j}")})
    schema = pa.schema(fields, metadata={"cases": json.dumps(meta), "reference": f"pyarrow {pa.__version__}", "seed": hex(SEED)})
    path = os.path.join(HERE, "reference_fixtures.arrow")
    with pa.OSFile(path, "wb") as f, pa.ipc.new_stream(f, schema) as w:
        w.write_batch(pa.record_batch(arrays, schema=schema))
    print(f"{len(rows)} cases, {len(pool)} distinct inputs -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
