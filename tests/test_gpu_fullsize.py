"""BASELINE.json's full sizes (1B rows; 2^30-1 keys for the sort) through size-independent
properties, checked on the device so nothing crosses PCIe: lengths = popcounts, checksums of
checksums (sum of the selected / gathered values), take round trips through a permutation and its
inverse, sortedness + permutation-ness of the sort indices, group-by totals.  torch is only the
checker here (random inputs, popcounts, sums); every operation under test goes through the C-ABI.

Set B200_FULLSIZE_ROWS to shrink the run (default 1_000_000_000)."""
import os

import numpy as np
import pyarrow as pa
import pytest

import arrow_b200.compute as bc
from arrow_b200 import DeviceArray

pytestmark = pytest.mark.gpu
N = int(os.environ.get("B200_FULLSIZE_ROWS", "1000000000"))
SEED = 0x0FF1CE


@pytest.fixture()
def torch_mod(ctx):
    import gc

    import torch
    gc.collect()
    torch.cuda.empty_cache()  # hand cached blocks back so the context's own pool can cudaMalloc them
    ctx.sync()
    ctx.trim()
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * N:
        pytest.skip(f"needs ~{70 * N >> 30} GiB of free device memory")
    return torch


def run(ctx, torch, fn):
    """torch (the checker) works on its own stream, the C-ABI calls on the context's: order them explicitly."""
    torch.cuda.synchronize()
    out = fn()
    ctx.sync()
    return out


def random_bitmap(torch, n, p_set, gen):
    """packed LSB-first bitmap with Bernoulli(p_set) bits, its bool expansion is produced chunk-wise"""
    out = torch.zeros((n + 7) // 8 + 64, dtype=torch.uint8, device="cuda")
    weights = (1 << torch.arange(8, device="cuda", dtype=torch.int32)).to(torch.uint8)
    chunk = 1 << 27
    count = 0
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        m8 = (m + 7) // 8 * 8
        v = torch.rand(m8, device="cuda", generator=gen) < p_set
        v[m:] = False
        count += int(v.sum().item())
        out[lo // 8: lo // 8 + m8 // 8] = (v.view(-1, 8).to(torch.uint8) * weights).sum(dim=1, dtype=torch.uint8)
    return out, count


def bits_to_bool(torch, bitmap, n):
    shifts = torch.arange(8, device="cuda", dtype=torch.uint8)
    return ((bitmap[: (n + 7) // 8, None] >> shifts) & 1).to(torch.bool).view(-1)[:n]


def as_tensor(torch, arr: DeviceArray, dtype, n=None):
    """zero-copy torch view of a DeviceArray's data buffer"""
    n = len(arr) if n is None else n
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Cai:  # __cuda_array_interface__ shim
        pass
    shim = _Cai()
    shim.__cuda_array_interface__ = {"shape": (n,), "typestr": {torch.int64: "<i8", torch.uint64: "<u8", torch.uint8: "|u1",
                                                              torch.float64: "<f8", torch.int32: "<i4"}[dtype],
                                     "data": (arr.buffers[1].ptr, False), "version": 2, "strides": (itemsize,)}
    return torch.as_tensor(shim, device="cuda")


def test_fullsize_filter(ctx, torch_mod):
    torch = torch_mod
    gen = torch.Generator(device="cuda").manual_seed(SEED)
    vals_t = torch.randint(-100, 101, (N,), dtype=torch.int64, device="cuda", generator=gen)
    valid_t, n_valid = random_bitmap(torch, N, 0.9, gen)
    mask_t, n_sel = random_bitmap(torch, N, 0.5, gen)
    vals = DeviceArray.from_pointers(ctx, pa.int64(), N, vals_t.data_ptr(), validity_ptr=valid_t.data_ptr(),
                                     null_count=N - n_valid)
    mask = DeviceArray.from_pointers(ctx, pa.bool_(), N, mask_t.data_ptr())
    out = run(ctx, torch, lambda: bc.filter(vals, mask))
    assert len(out) == n_sel
    sel = bits_to_bool(torch, mask_t, N)
    ok = bits_to_bool(torch, valid_t, N)
    assert out.null_count == n_sel - int((sel & ok).sum().item())
    got = as_tensor(torch, out, torch.int64)
    # order-sensitive checksum: sum(value * (position mod 1021)) over the compacted output vs the reference positions
    pos = torch.cumsum(sel, 0, dtype=torch.int64) - 1
    want = (vals_t * (pos % 1021))[sel].sum().item()
    have = (got * (torch.arange(n_sel, device="cuda") % 1021)).sum().item()
    assert have == want
    out_ok = bits_to_bool(torch, as_tensor(torch, _validity_view(out), torch.uint8, (n_sel + 7) // 8), n_sel)
    assert torch.equal(out_ok, ok[sel])


def _validity_view(arr: DeviceArray):
    class _V:
        pass
    v = _V()
    v.buffers = [None, arr.buffers[0]]
    return v


def test_fullsize_take_round_trip(ctx, torch_mod):
    torch = torch_mod
    gen = torch.Generator(device="cuda").manual_seed(SEED + 1)
    vals_t = torch.rand(N, dtype=torch.float64, device="cuda", generator=gen) * 1e6
    perm_t = torch.randperm(N, device="cuda", generator=gen)
    inv_t = torch.empty_like(perm_t)
    inv_t[perm_t] = torch.arange(N, device="cuda")
    vals = DeviceArray.from_pointers(ctx, pa.float64(), N, vals_t.data_ptr())
    perm = DeviceArray.from_pointers(ctx, pa.int64(), N, perm_t.data_ptr())
    inv = DeviceArray.from_pointers(ctx, pa.int64(), N, inv_t.data_ptr())
    once = run(ctx, torch, lambda: bc.take(vals, perm))
    assert as_tensor(torch, once, torch.float64)[12345].item() == vals_t[perm_t[12345]].item()
    back = run(ctx, torch, lambda: bc.take(once, inv))
    assert torch.equal(as_tensor(torch, back, torch.float64), vals_t)


def test_fullsize_take_with_validity_bands(ctx, torch_mod):
    """values carry a validity bitmap (125 MB at 1B rows, probed in bands): out validity = values' validity at the index,
    checked on the device against the bool expansion; the gathered values are checked through the inverse permutation."""
    torch = torch_mod
    gen = torch.Generator(device="cuda").manual_seed(SEED + 11)
    vals_t = torch.randint(-2**40, 2**40, (N,), dtype=torch.int64, device="cuda", generator=gen)
    valid_t, n_valid = random_bitmap(torch, N, 0.9, gen)
    idx_t = torch.randint(0, N, (N,), dtype=torch.int64, device="cuda", generator=gen)
    vals = DeviceArray.from_pointers(ctx, pa.int64(), N, vals_t.data_ptr(), validity_ptr=valid_t.data_ptr(), null_count=N - n_valid)
    idx = DeviceArray.from_pointers(ctx, pa.int64(), N, idx_t.data_ptr())
    out = run(ctx, torch, lambda: bc.take(vals, idx))
    ok = bits_to_bool(torch, valid_t, N)
    want_valid = ok[idx_t]
    assert out.null_count == N - int(want_valid.sum().item())
    out_ok = bits_to_bool(torch, as_tensor(torch, _validity_view(out), torch.uint8, (N + 7) // 8), N)
    assert torch.equal(out_ok, want_valid)
    del out_ok, want_valid, ok
    got = as_tensor(torch, out, torch.int64)
    assert int((got - vals_t[idx_t]).abs().max().item()) == 0


@pytest.mark.parametrize("beyond_2_30", [False, True])
def test_fullsize_sort_indices(ctx, torch_mod, beyond_2_30):
    """beyond_2_30: 2^30 + 70001 rows -- the 64-bit look-back cells of the radix passes (round 1 refused >= 2^30 rows)"""
    torch = torch_mod
    if beyond_2_30 and N < 1_000_000_000:
        pytest.skip("full-size run only")
    n = (1 << 30) + 70001 if beyond_2_30 else min(N, (1 << 30) - 1)
    gen = torch.Generator(device="cuda").manual_seed(SEED + 2)
    keys_t = torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device="cuda", generator=gen)
    valid_t, n_valid = random_bitmap(torch, n, 0.9, gen)
    keys = DeviceArray.from_pointers(ctx, pa.int64(), n, keys_t.data_ptr(), validity_ptr=valid_t.data_ptr(),
                                     null_count=n - n_valid)
    idx = as_tensor(torch, run(ctx, torch, lambda: bc.array_sort_indices(keys)), torch.int64)
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    seen[idx] = True
    assert bool(seen.all())  # a permutation
    ok = bits_to_bool(torch, valid_t, n)
    assert bool(ok[idx[:n_valid]].all()) and not bool(ok[idx[n_valid:]].any())  # nulls at the end
    sk = keys_t[idx[:n_valid]]
    d = sk[1:] - sk[:-1]
    assert bool((sk[1:] >= sk[:-1]).all())
    ties = d == 0
    assert bool((idx[1:n_valid][ties] > idx[:n_valid - 1][ties]).all())  # stable
    assert bool((idx[n_valid + 1:] > idx[n_valid:-1]).all())  # nulls keep index order


def test_fullsize_group_by(ctx, torch_mod):
    torch = torch_mod
    groups = 10_000_000
    gen = torch.Generator(device="cuda").manual_seed(SEED + 3)
    keys_t = torch.randint(0, groups, (N,), dtype=torch.int64, device="cuda", generator=gen)
    vals_t = torch.randint(-100, 101, (N,), dtype=torch.int64, device="cuda", generator=gen)
    valid_t, n_valid = random_bitmap(torch, N, 0.9, gen)
    keys = DeviceArray.from_pointers(ctx, pa.int64(), N, keys_t.data_ptr())
    vals = DeviceArray.from_pointers(ctx, pa.int64(), N, vals_t.data_ptr(), validity_ptr=valid_t.data_ptr(),
                                     null_count=N - n_valid)
    g = bc.GroupBySumCount(pa.int64(), pa.int64(), expected_groups=groups, ctx=ctx)
    run(ctx, torch, lambda: g.consume(keys, vals))
    out_keys, sums, counts = run(ctx, torch, g.finalize)
    ok = bits_to_bool(torch, valid_t, N)
    want_sum = torch.zeros(groups, dtype=torch.int64, device="cuda").index_add_(0, keys_t[ok], vals_t[ok])
    want_cnt = torch.bincount(keys_t[ok], minlength=groups)
    present = torch.bincount(keys_t, minlength=groups) > 0
    k = as_tensor(torch, out_keys, torch.int64)
    assert len(out_keys) == int(present.sum().item())
    assert bool(present[k].all()) and int(torch.unique(k).numel()) == len(out_keys)
    assert torch.equal(as_tensor(torch, counts, torch.int64), want_cnt[k])
    s = as_tensor(torch, sums, torch.int64)
    has = want_cnt[k] > 0  # groups whose values are all null finalize to a null sum (min_count = 1)
    assert torch.equal(s[has], want_sum[k][has])
    assert sums.null_count == int((~has).sum().item())


def test_dense_group_by_hot_key_drains(ctx, torch_mod):
    """direct-addressed group-by with a HOT key: more than 2^26 rows of one key force the guarded drain between
    sub-batches (groupby_dense.cuh: the packed count field of a sub-batch holds 2^26 rows for this value window)"""
    torch = torch_mod
    n = min(N, 300_000_000)
    groups = 1_000_000
    gen = torch.Generator(device="cuda").manual_seed(SEED + 9)
    keys_t = torch.randint(0, groups, (n,), dtype=torch.int64, device="cuda", generator=gen)
    keys_t[torch.rand(n, device="cuda", generator=gen) < 0.6] = 777            # ~180M rows of one key
    vals_t = torch.randint(-100, 101, (n,), dtype=torch.int64, device="cuda", generator=gen)
    valid_t, n_valid = random_bitmap(torch, n, 0.9, gen)
    keys = DeviceArray.from_pointers(ctx, pa.int64(), n, keys_t.data_ptr())
    vals = DeviceArray.from_pointers(ctx, pa.int64(), n, vals_t.data_ptr(), validity_ptr=valid_t.data_ptr(), null_count=n - n_valid)
    g = bc.GroupBySumCount(pa.int64(), pa.int64(), ctx=ctx)
    run(ctx, torch, lambda: g.consume(keys, vals))
    assert g.path_counts()["dense"] == 1
    out_keys, sums, counts = run(ctx, torch, g.finalize)
    ok = bits_to_bool(torch, valid_t, n)
    want_sum = torch.zeros(groups, dtype=torch.int64, device="cuda").index_add_(0, keys_t[ok], vals_t[ok])
    want_cnt = torch.bincount(keys_t[ok], minlength=groups)
    present = torch.bincount(keys_t, minlength=groups) > 0
    k = as_tensor(torch, out_keys, torch.int64)
    assert len(out_keys) == int(present.sum().item())
    assert torch.equal(as_tensor(torch, counts, torch.int64), want_cnt[k])
    has = want_cnt[k] > 0
    assert torch.equal(as_tensor(torch, sums, torch.int64)[has], want_sum[k][has])
    assert int(want_cnt[777].item()) > (1 << 26)
