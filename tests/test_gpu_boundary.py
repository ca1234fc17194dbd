"""Drop-in boundary on the device: the caller-supplied allocator (b2_context_set_allocator -- how the host keeps
every output inside ITS MemoryManager pool) and the event entry points."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


def test_outputs_come_from_the_installed_allocator():
    import torch

    import arrow_b200.compute as bc
    from arrow_b200 import DeviceArray, _cabi as cabi
    from arrow_b200.device import Context, check
    from oracle import arrow_oracle as ora

    ctx = Context(0)  # a private context: the shared one keeps its built-in pool
    live, served = {}, []

    @C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)
    def alloc(nbytes, stream, user):
        t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda")  # the "host-side pool" of this test
        live[t.data_ptr()] = t
        served.append((t.data_ptr(), int(nbytes)))
        return t.data_ptr()

    @C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)
    def free(ptr, nbytes, stream, user):
        torch.cuda.synchronize()   # stream-ordered reuse is the pool's business; this toy pool just waits
        assert ptr in live, "free of a pointer the allocator never handed out"
        del live[ptr]

    check(ctx.lib.b2_context_set_allocator(ctx.handle, C.cast(alloc, C.c_void_p), C.cast(free, C.c_void_p), None))
    rng = np.random.default_rng(0x0FF1CE)
    n = 100_003
    values = pa.array(rng.uniform(0, 1e6, n), mask=rng.random(n) < 0.1)
    idx = pa.array(rng.integers(0, n, n, dtype=np.int64))
    mask = pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.05)
    keys = pa.array(rng.integers(0, 500, n, dtype=np.int64))
    dv, di, dm, dk = (DeviceArray.from_arrow(x, ctx) for x in (values, idx, mask, keys))
    inputs = {b.ptr for a in (dv, di, dm, dk) for b in a.buffers if b is not None}
    assert inputs <= set(live), "inputs uploaded through the context must come from the allocator too"

    outs = {
        "take": bc.take(dv, di), "filter": bc.filter(dv, dm), "cast": bc.cast(dv, pa.float32(), safe=False),
        "sort": bc.array_sort_indices(dv), "cmp": bc.greater(dv, 5e5), "and": bc.and_kleene(dm, dm),
    }
    outs["add"] = bc.add(outs["cast"], outs["cast"])
    (uk,), (s, c) = bc.group_by([dk], [("hash_sum", di, None), ("hash_count", di, None)], fused=False)
    outs.update(uk=uk, s=s, c=c)
    for name, arr in outs.items():
        for b in arr.buffers:
            if b is not None:
                assert b.ptr in live, f"{name}: output buffer {hex(b.ptr)} was not allocated through the callbacks"
    assert outs["take"].to_arrow().equals(ora.take(values, idx))
    assert outs["filter"].to_arrow().equals(ora.filter(values, mask))
    assert len(served) > len(outs)           # temporaries were served by the callbacks as well ...
    del outs, dv, di, dm, dk, uk, s, c, arr, b, name
    import gc
    gc.collect()
    assert not live, f"{len(live)} allocations were never returned to the allocator (temporaries must be freed)"
    ctx.lib.b2_context_destroy(ctx.handle)
    ctx.handle = None


def test_event_entry_points(ctx):
    from arrow_b200.device import check
    ev = C.c_void_p()
    check(ctx.lib.b2_event_create(ctx.handle, C.byref(ev)))
    check(ctx.lib.b2_event_record(ctx.handle, ev, ctx.stream))
    check(ctx.lib.b2_stream_wait_event(ctx.handle, ctx.stream, ev))
    check(ctx.lib.b2_event_synchronize(ev))
    check(ctx.lib.b2_event_destroy(ev))
