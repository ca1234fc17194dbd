#!/usr/bin/env python
"""Does handing evict_last L2 lines back (launch_l2_demote, B2_L2_DEMOTE=1) help or hurt the kernels that follow?
One process per setting (the library reads the knob once): fused dense group-by (config 3) timed before and after a
banded Take, and the cast + add that follow a take.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import torch

import arrow_b200.compute as bc
from arrow_b200 import Context, DeviceArray
from bench import SEED, make_validity

n = int(os.environ.get("ROWS", "1000000000"))
torch.cuda.set_device(0)
ctx = Context.get(0)
ctx.stream = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device="cuda")
gen.manual_seed(SEED)
keys_t = torch.randint(0, 10_000_000, (n,), dtype=torch.int64, device="cuda", generator=gen)
vals_t = torch.randint(-100, 101, (n,), dtype=torch.int64, device="cuda", generator=gen)
vvalid_t, v_nulls = make_validity(torch, n, gen)
keys = DeviceArray.from_pointers(ctx, pa.int64(), n, keys_t.data_ptr())
vals = DeviceArray.from_pointers(ctx, pa.int64(), n, vals_t.data_ptr(), validity_ptr=vvalid_t.data_ptr(), null_count=v_nulls)
fv_t = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
fvalid_t, f_nulls = make_validity(torch, n, gen)
idx_t = torch.randint(0, n, (n,), dtype=torch.int64, device="cuda", generator=gen)
fvals = DeviceArray.from_pointers(ctx, pa.float64(), n, fv_t.data_ptr(), validity_ptr=fvalid_t.data_ptr(), null_count=f_nulls)
idx = DeviceArray.from_pointers(ctx, pa.int64(), n, idx_t.data_ptr())
other_t = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen)
other = DeviceArray.from_pointers(ctx, pa.float32(), n, other_t.data_ptr())
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


def group_by_ms(reps=3):
    out = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record()
        g = bc.GroupBySumCount(pa.int64(), pa.int64(), expected_groups=10_000_000, ctx=ctx)
        g.consume(keys, vals)
        r = g.finalize()
        b.record()
        b.synchronize()
        out.append(round(a.elapsed_time(b), 3))
        del r, g
    return out


def take_then_tail():
    a, b, c = ev(), ev(), ev()
    a.record()
    t = bc.take(fvals, idx)
    b.record()
    x = bc.cast(t, pa.float32(), safe=False)
    y = bc.add(x, other)
    c.record()
    c.synchronize()
    del t, x, y
    return round(a.elapsed_time(b), 3), round(b.elapsed_time(c), 3)


res = {"B2_L2_DEMOTE": os.environ.get("B2_L2_DEMOTE", "unset"), "rows": n}
group_by_ms(1)
take_then_tail()
res["group_by_ms_first"] = group_by_ms()
res["take_ms__cast_add_ms"] = [take_then_tail() for _ in range(3)]
res["group_by_ms_after_takes"] = group_by_ms()
print(json.dumps(res), flush=True)
