#!/usr/bin/env python
"""Two Take launches for an ncu DRAM-traffic capture at the BENCHMARKED size (separate
value traffic from validity-bitmap traffic instead of extrapolating a 200M-row capture):
  launch 1: take(float64 values WITH a validity bitmap, random int64 indices)  -> take_kernel<8,long,true>
  launch 2: take(the same values WITHOUT validity, same indices)               -> take_kernel<8,long,false>
Run under:  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
            --clock-control none -k regex:take_kernel -c 2 --csv --log-file gpurun_out/take_traffic_r02.csv python scripts/take_traffic.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import torch

import arrow_b200.compute as bc
from arrow_b200 import Context, DeviceArray
from bench import SEED, make_validity

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000_000)
ap.add_argument("--variants", default="valid,novalid,fused")
args = ap.parse_args()
n = args.rows
torch.cuda.set_device(0)
ctx = Context.get(0)
gen = torch.Generator(device="cuda")
gen.manual_seed(SEED)
values_t = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 1e6
vvalid_t, v_nulls = make_validity(torch, n, gen)
idx_t = torch.randint(0, n, (n,), dtype=torch.int64, device="cuda", generator=gen)
torch.cuda.synchronize()
idx = DeviceArray.from_pointers(ctx, pa.int64(), n, idx_t.data_ptr())
for v in args.variants.split(","):
    if v == "fused":   # launch 3: the fused take+cast+add kernel of the bench step (take_cast_arith_kernel<double,long,float,true>)
        other_t = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen) * 1e6
        ovalid_t, o_nulls = make_validity(torch, n, gen)
        torch.cuda.synchronize()
        values = DeviceArray.from_pointers(ctx, pa.float64(), n, values_t.data_ptr(), validity_ptr=vvalid_t.data_ptr(), null_count=v_nulls)
        other = DeviceArray.from_pointers(ctx, pa.float32(), n, other_t.data_ptr(), validity_ptr=ovalid_t.data_ptr(), null_count=o_nulls)
        out = bc.take_cast_arith(values, idx, pa.float32(), "add", other)
        ctx.sync()
        print(v, "nulls", out.null_count, flush=True)
        del out
        continue
    if v == "valid":
        values = DeviceArray.from_pointers(ctx, pa.float64(), n, values_t.data_ptr(), validity_ptr=vvalid_t.data_ptr(), null_count=v_nulls)
    else:
        values = DeviceArray.from_pointers(ctx, pa.float64(), n, values_t.data_ptr())
    out = bc.take(values, idx)
    ctx.sync()
    print(v, "nulls", out.null_count, flush=True)
    del out
