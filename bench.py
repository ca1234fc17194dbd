#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1]: Take + Cast(float64->float32) + Add over 1B rows per
GPU, null_probability 0.1, as three CallFunction-equivalent calls through the C-ABI:

    out = add(cast(take(values, indices), float32), other)

One "step" = one pass of that pipeline over one synthetic batch.
  value : rows/s, inputs already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e   : rows/s through the same public calls starting from pinned HOST buffers, H2D of all
          inputs and D2H of the result column inside the timed region
  roofline : the dominant kernel (take_kernel, random int64 indices into float64), algorithmic
          bytes (24.25 B/row) / its CUDA-event time, against the measured copy peak
  cpu_baseline / --impl reference : the reference binary (pyarrow 24.0.0 = libarrow_compute.so.2400,
          the same kernels as /root/reference for this path) on the host cores, on a bounded sample.
Multi-GPU: rows shard by range, `values` replicated per GPU, no collective on the
data path => weak scaling, one process per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_TAKE, ALG_CAST, ALG_ADD = 24.25, 12.25, 12.375  # bytes/row, SURVEY section 8d
ALG_PIPELINE = ALG_TAKE + ALG_CAST + ALG_ADD          # 48.875
NULL_P = 0.1
SEED = 0x0FF1CE


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        # NVML in-process (same counters nvidia-smi prints, but every 10 ms: the device-resident timed region
        # is < 200 ms long); the nvidia-smi subprocess loop is the fallback
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def _nvml_loop(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [0x8, 0x40, 0x20, 0x4]  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap (nvml.h)
        while not self.stop_flag.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = int(reasons(h))
            self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            self.stop_flag.wait(0.01)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU kernels on the host cores
# ------------------------------------------------------------------------------------------------
def reference_pipeline_rate(sample_rows, steps, warmup, threads):
    """rows/s of add(cast(take(values, idx), f32), other) with pyarrow.compute, `threads` row-range
    slices in flight (CallFunction itself is single-threaded, compute/exec.h:85-91)."""
    import concurrent.futures as cf

    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc

    rng = np.random.default_rng(SEED)
    n = sample_rows
    values = pa.array(rng.uniform(0, 1e6, n), pa.float64(), mask=rng.random(n) < NULL_P)
    indices = pa.array(rng.integers(0, n, n, dtype=np.int64))
    other = pa.array(rng.uniform(0, 1e6, n).astype(np.float32), pa.float32(), mask=rng.random(n) < NULL_P)
    pa.set_cpu_count(threads)
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)

    def part(i):
        lo, hi = int(bounds[i]), int(bounds[i + 1])
        t = pc.take(values, indices.slice(lo, hi - lo))
        c = pc.cast(t, pa.float32(), safe=False)
        return pc.add(c, other.slice(lo, hi - lo))

    pool = cf.ThreadPoolExecutor(threads)
    for _ in range(warmup):
        list(pool.map(part, range(threads)))
    t0 = time.perf_counter()
    for _ in range(steps):
        list(pool.map(part, range(threads)))
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_rows
    rate, per_step = reference_pipeline_rate(sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "rows/sec", "value": rate, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64->f32", "data": "synthetic",
        "config": {"workload": "Take(float64,int64 idx)+Cast(float64->float32)+Add(float32), null_probability=0.1",
                   "rows_per_step": sample, "note": "bounded sample of the 1B-row workload on host cores"},
        "cpu_baseline": {"value": rate, "unit": "rows/s", "cores": threads, "kind": "reference",
                         "sample": f"{sample} rows/step x {args.steps} steps, pyarrow 24.0.0 libarrow_compute, {threads} row-range threads"},
        "e2e": {"value": rate, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def pack_bits(torch, valid):
    """bool[n] (n % 8 == 0) -> LSB-first bitmap bytes"""
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=valid.device)
    return (valid.view(-1, 8).to(torch.uint8) * w).sum(dim=1, dtype=torch.uint8)


def make_validity(torch, n, gen):
    out = torch.empty((n + 7) // 8 + 64, dtype=torch.uint8, device="cuda")
    out.zero_()
    chunk = 1 << 27
    nulls = 0
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        m8 = (m + 7) // 8 * 8
        v = torch.rand(m8, device="cuda", generator=gen) >= NULL_P
        if m8 != m:
            v[m:] = False
        nulls += int(m - v[:m].sum().item())
        out[lo // 8: lo // 8 + m8 // 8] = pack_bits(torch, v)
        del v
    return out, nulls


def run_gpu(args):
    import numpy as np
    import pyarrow as pa
    import torch
    import torch.distributed as dist

    import arrow_b200.compute as bc
    from arrow_b200 import Context, DeviceArray, PinnedBuffer, _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context.get(local)
    # a real (non-NULL) stream: the C-ABI treats stream 0 as "use the context's own stream", and
    # torch.cuda.Event only times the stream it is recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.stream = stream.cuda_stream  # all C-ABI calls are ordered on torch's current stream
    lib = _cabi.lib()

    n = args.rows
    gen = torch.Generator(device="cuda")
    gen.manual_seed(SEED + rank)
    # ---- synthetic inputs, resident in HBM ----
    values_t = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 1e6
    vvalid_t, v_nulls = make_validity(torch, n, gen)
    idx_t = torch.randint(0, n, (n,), dtype=torch.int64, device="cuda", generator=gen)
    other_t = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen) * 1e6
    ovalid_t, o_nulls = make_validity(torch, n, gen)
    values = DeviceArray.from_pointers(ctx, pa.float64(), n, values_t.data_ptr(), validity_ptr=vvalid_t.data_ptr(), null_count=v_nulls)
    indices = DeviceArray.from_pointers(ctx, pa.int64(), n, idx_t.data_ptr())
    other = DeviceArray.from_pointers(ctx, pa.float32(), n, other_t.data_ptr(), validity_ptr=ovalid_t.data_ptr(), null_count=o_nulls)

    def pipeline(v, i, o):
        t = bc.take(v, i)
        c = bc.cast(t, pa.float32(), safe=False)
        return bc.add(c, o)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also fills the pool so the timed region never calls cudaMalloc) ----
    for _ in range(args.warmup):
        out = pipeline(values, indices, other)
    out_nulls = out.null_count
    del out

    # ---- per-kernel timing (events around each call; the call = its kernel + a bitmap kernel) ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k_ms = np.zeros(3)
    sync_all()
    for _ in range(args.steps):
        ev[0].record(stream)
        t = bc.take(values, indices)
        ev[1].record(stream)
        c = bc.cast(t, pa.float32(), safe=False)
        ev[2].record(stream)
        o = bc.add(c, other)
        ev[3].record(stream)
        torch.cuda.synchronize()
        k_ms += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
        del t, c, o
    k_ms /= args.steps

    # ---- timed region: K steps, device resident ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.b2_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    start.record(stream)
    for _ in range(args.steps):
        out = pipeline(values, indices, other)
        del out
    stop.record(stream)
    sync_all()
    total_ms = max_over_ranks(start.elapsed_time(stop))
    launches = lib.b2_launch_count() - launches0
    clocks = sampler.summary() if sampler else None
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- sorted-index variant of take (SURVEY section 8d asks for both) ----
    inc = torch.randint(0, 3, (n,), dtype=torch.int64, device="cuda", generator=gen)
    sorted_idx_t = torch.clamp(torch.cumsum(inc, 0), max=n - 1)
    del inc
    sorted_idx = DeviceArray.from_pointers(ctx, pa.int64(), n, sorted_idx_t.data_ptr())
    t = bc.take(values, sorted_idx)
    del t
    torch.cuda.synchronize()
    ev[0].record(stream)
    for _ in range(args.steps):
        t = bc.take(values, sorted_idx)
        del t
    ev[1].record(stream)
    torch.cuda.synchronize()
    take_sorted_ms = ev[0].elapsed_time(ev[1]) / args.steps
    del sorted_idx, sorted_idx_t

    # ---- e2e: pinned host inputs -> H2D -> pipeline -> D2H of the result column ----
    e2e_rows = min(n, args.e2e_rows) if args.e2e_rows else n
    try:
        import psutil
        avail = psutil.virtual_memory().available
        need = e2e_rows * 24.5 * (world if world > 1 else 1)
        if need > 0.5 * avail:
            e2e_rows = int(0.5 * avail / (24.5 * world)) // 64 * 64
    except Exception:
        pass
    m = e2e_rows
    bm = (m + 7) // 8
    h_values, h_vvalid = PinnedBuffer(8 * m), PinnedBuffer(bm)
    h_idx, h_other, h_ovalid = PinnedBuffer(8 * m), PinnedBuffer(4 * m), PinnedBuffer(bm)
    h_out, h_outvalid = PinnedBuffer(4 * m), PinnedBuffer(bm + 8)
    idx_small = idx_t[:m] % m if m != n else idx_t
    for hb, src in ((h_values, values_t[:m]), (h_vvalid, vvalid_t[:bm]), (h_idx, idx_small), (h_other, other_t[:m]), (h_ovalid, ovalid_t[:bm])):
        ctx.d2h(hb.ptr, src.data_ptr(), hb.size)
    ctx.sync()
    del idx_small
    d_values, d_vvalid = ctx.alloc(8 * m), ctx.alloc(bm + 64)
    d_idx, d_other, d_ovalid = ctx.alloc(8 * m), ctx.alloc(4 * m), ctx.alloc(bm + 64)
    h2d_bytes = 8 * m + bm + 8 * m + 4 * m + bm
    d2h_bytes = 4 * m + bm

    # The e2e leg is chunked the way the reference chunks: `indices` / `other` travel as K chunks (a
    # ChunkedArray argument to take/add, each chunk one CallFunction-equivalent call), so the H2D copy of
    # chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 overlap on three streams.  `values`
    # (the gathered column) must be fully resident before the first chunk is gathered.
    K = max(1, args.e2e_chunks)
    copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
    bounds = [min(m, (m * k // K) // 64 * 64) for k in range(K)] + [m]

    def e2e_step():
        ci, co, cs = copy_in.cuda_stream, copy_out.cuda_stream, stream.cuda_stream
        ctx.h2d(d_values.ptr, h_values.ptr, h_values.size, ci)
        ctx.h2d(d_vvalid.ptr, h_vvalid.ptr, h_vvalid.size, ci)
        ev_vals = torch.cuda.Event()
        ev_vals.record(copy_in)
        ready = []
        for k in range(K):
            lo, hi = bounds[k], bounds[k + 1]
            ctx.h2d(d_idx.ptr + 8 * lo, h_idx.ptr + 8 * lo, 8 * (hi - lo), ci)
            ctx.h2d(d_other.ptr + 4 * lo, h_other.ptr + 4 * lo, 4 * (hi - lo), ci)
            ctx.h2d(d_ovalid.ptr + lo // 8, h_ovalid.ptr + lo // 8, (hi - lo + 7) // 8, ci)
            e = torch.cuda.Event()
            e.record(copy_in)
            ready.append(e)
        v = DeviceArray.from_pointers(ctx, pa.float64(), m, d_values.ptr, validity_ptr=d_vvalid.ptr, null_count=-1)
        stream.wait_event(ev_vals)
        keep, nulls = [], 0
        for k in range(K):
            lo, hi = bounds[k], bounds[k + 1]
            stream.wait_event(ready[k])
            i = DeviceArray.from_pointers(ctx, pa.int64(), hi - lo, d_idx.ptr + 8 * lo)
            o = DeviceArray.from_pointers(ctx, pa.float32(), hi - lo, d_other.ptr + 4 * lo, validity_ptr=d_ovalid.ptr + lo // 8,
                                          null_count=-1)
            r = pipeline(v, i, o)          # three C-ABI calls on `stream`
            done = torch.cuda.Event()
            done.record(stream)
            copy_out.wait_event(done)
            ctx.d2h(h_out.ptr + 4 * lo, r.buffers[1].ptr, 4 * (hi - lo), co)
            if r.buffers[0] is not None:   # chunk validity lands at its own (byte aligned: lo % 64 == 0) position
                ctx.d2h(h_outvalid.ptr + lo // 8, r.buffers[0].ptr, (hi - lo + 7) // 8, co)
            keep.append(r)                 # buffers stay alive until the D2H stream drains
            nulls += r.null_count
        copy_out.synchronize()
        return nulls

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    start.record(stream)
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    stop.record(stream)
    sync_all()
    e2e_ms = max_over_ranks(max(start.elapsed_time(stop), (time.perf_counter() - t0) * 1e3)) / e2e_steps
    e2e_value = m * world / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    take_gbs = ALG_TAKE * n / (k_ms[0] * 1e-3) / 1e9
    kernels = [
        {"name": "take_kernel<8,int64> (random idx)", "ms": float(k_ms[0]), "alg_bytes_per_row": ALG_TAKE, "gbs": take_gbs, "frac": take_gbs / peak},
        {"name": "take_kernel<8,int64> (monotonic idx)", "ms": float(take_sorted_ms), "alg_bytes_per_row": ALG_TAKE,
         "gbs": ALG_TAKE * n / (take_sorted_ms * 1e-3) / 1e9, "frac": ALG_TAKE * n / (take_sorted_ms * 1e-3) / 1e9 / peak},
        {"name": "map1_kernel<double,float> (cast)", "ms": float(k_ms[1]), "alg_bytes_per_row": ALG_CAST,
         "gbs": ALG_CAST * n / (k_ms[1] * 1e-3) / 1e9, "frac": ALG_CAST * n / (k_ms[1] * 1e-3) / 1e9 / peak},
        {"name": "map2_kernel<float> (add)", "ms": float(k_ms[2]), "alg_bytes_per_row": ALG_ADD,
         "gbs": ALG_ADD * n / (k_ms[2] * 1e-3) / 1e9, "frac": ALG_ADD * n / (k_ms[2] * 1e-3) / 1e9 / peak},
    ]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "take_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_row"] * n
        except Exception:
            traffic = None
    cores = os.cpu_count() or 1
    cpu_rate, _ = reference_pipeline_rate(args.cpu_sample_rows, 2, 1, cores)
    line = {
        "metric": "rows/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64->f32", "data": "synthetic",
        "config": {"workload": "Take(float64,int64 idx)+Cast(float64->float32)+Add(float32), null_probability=0.1, 1B rows/GPU"
                   if n == 1_000_000_000 else f"Take+Cast(f64->f32)+Add, null_probability=0.1, {n} rows/GPU",
                   "rows_per_gpu": n, "indices": "uniform random int64", "l2": "inputs (20 GB) are far larger than the 126 MB L2",
                   "pipeline_alg_bytes_per_row": ALG_PIPELINE, "pipeline_gbs": ALG_PIPELINE * n * world / (ms_per_step * 1e-3) / 1e9,
                   "out_null_count": int(out_nulls)},
        "roofline": {"bound": "hbm", "achieved": take_gbs, "peak": peak, "unit": "GB/s", "frac": take_gbs / peak,
                     "traffic": traffic, "kernel": "take_kernel<8,int64_t,true>", "peak_source": peak_src,
                     "alg_bytes_per_launch": ALG_TAKE * n},
        "kernels": kernels,
        "cpu_baseline": {"value": cpu_rate, "unit": "rows/s", "cores": cores, "kind": "reference",
                         "sample": f"{args.cpu_sample_rows} rows x 2 steps of the same pipeline, pyarrow 24.0.0 libarrow_compute, {cores} row-range threads"},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "rows_per_gpu": m, "ms_per_step": e2e_ms, "chunks": K},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows per GPU for the host-buffer leg (0 = same as --rows)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="chunks of the indices/other columns in the host-buffer leg")
    ap.add_argument("--cpu-sample-rows", type=int, default=1 << 25)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
