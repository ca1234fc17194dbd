"""Launched by tests/test_gpu_distributed.py under torch.distributed.run (one rank per GPU, NCCL):
runs arrow_b200.distributed on device shards and checks the result against the single-process oracle."""
import os
import sys

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from arrow_b200 import Context, DeviceArray
    from arrow_b200 import distributed as d
    from oracle import arrow_oracle as ora
    ctx = Context.get(local)
    ops = d.DeviceOps(ctx)
    n = int(os.environ.get("B2_DIST_ROWS", "400003"))
    rng = np.random.default_rng(0x0FF1CE)
    keys = pa.array(rng.integers(0, 5000, n, dtype=np.int64), mask=rng.random(n) < 0.03)
    vals = pa.array(rng.integers(-100, 100, n, dtype=np.int64), mask=rng.random(n) < 0.1)
    sortkeys = pa.array(rng.integers(-2**62, 2**62, n, dtype=np.int64), mask=rng.random(n) < 0.1)
    lo, hi = rank * n // world, (rank + 1) * n // world
    dk, dv, ds = (DeviceArray.from_arrow(x.slice(lo, hi - lo), ctx) for x in (keys, vals, sortkeys))

    fvals = pa.array(rng.uniform(-5, 5, n), mask=rng.random(n) < 0.1)
    dfv = DeviceArray.from_arrow(fvals.slice(lo, hi - lo), ctx)
    uniq, (os_, oc) = ora.group_by([keys], [("hash_sum", vals, None), ("hash_count", vals, None)])
    want = sorted(zip(uniq[0].to_pylist(), os_.to_pylist(), oc.to_pylist()), key=lambda t: (t[0] is None, t[0]))
    funiq, (fs, fc) = ora.group_by([keys], [("hash_sum", fvals, None), ("hash_count", fvals, None)])
    fwant = sorted(zip(funiq[0].to_pylist(), fs.to_pylist(), fc.to_pylist()), key=lambda t: (t[0] is None, t[0]))
    want_sorted = ora.sort_indices(sortkeys).to_pylist()
    # both transports: the C-ABI NCCL path (b2_comm_*, what bench.py uses) and torch.distributed
    for name, xchg in (("b2_comm", d.B2CommExchange(ctx)), ("torch", d.TorchExchange())):
        k, s, c = d.group_by_sum_count(dk, dv, ops, xchg)
        mine = list(zip(k.to_arrow().to_pylist(), s.to_arrow().to_pylist(), c.to_arrow().to_pylist()))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        k, s, c = d.group_by_sum_count(dk, dfv, ops, xchg)
        fmine = list(zip(k.to_arrow().to_pylist(), s.to_arrow().to_pylist(), c.to_arrow().to_pylist()))
        fgathered = [None] * world
        dist.all_gather_object(fgathered, fmine)
        seg, nulls = d.sort_indices(ds, ops, xchg)
        segs = [None] * world
        dist.all_gather_object(segs, (seg.cpu().tolist(), nulls.cpu().tolist()))
        if rank == 0:
            got = sorted([t for g in gathered for t in g], key=lambda t: (t[0] is None, t[0]))
            assert got == want, f"{name}: distributed group-by mismatch"
            fgot = sorted([t for g in fgathered for t in g], key=lambda t: (t[0] is None, t[0]))
            assert [(a, c_) for a, _, c_ in fgot] == [(a, c_) for a, _, c_ in fwant], f"{name}: float group-by keys/counts"
            for (_, x, _), (_, y, _) in zip(fgot, fwant):
                assert (x is None and y is None) or abs(x - y) <= 1e-9 * max(1.0, abs(y)), f"{name}: float sums"
            cat = [i for sg in segs for i in sg[0]] + [i for sg in segs for i in sg[1]]
            assert cat == want_sorted, f"{name}: distributed sort mismatch"
            print(f"DIST OK transport={name} world={world} rows={n} groups={len(want)} segment_sizes={[len(sg[0]) for sg in segs]}", flush=True)
        if hasattr(xchg, "close"):
            xchg.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
