"""Multi-GPU execution of the two hot-path operators that need an exchange step.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU for the tests).
Filter / Cast / arithmetic / compare / Take shard by row range with no collective (bench.py does
exactly that); hash-aggregate and SortIndices do one local pass, ONE all-to-all, and one local
merge -- the shape of the reference's own per-thread scheme
(GroupByNode: per-thread Grouper + aggregators, then Merge by re-consuming uniques,
acero/groupby_aggregate_node.cc:211-218,255-298), with GPUs in place of threads:

  group_by_sum_count : local fused group-by -> partial groups (key, sum, count) ->
                       destination = hash(key) % P (b2_hash_partition) -> stable partition
                       (sort_indices on the destination id + take) -> all-to-all-v ->
                       owner merges partials with Grouper + hash_sum (sum of sums, sum of counts).
  sort_indices       : sample -> all_gather -> P-1 splitters -> destination = range id
                       (b2_range_partition) -> stable partition of (key, global row) ->
                       all-to-all-v -> local stable sort; nulls never move.
                       Result = rank-ordered concatenation of the value segments, then of the
                       null segments (AtEnd) -- identical to the single-GPU answer.

The algorithms are written against a small `ops` interface so the same code runs on device
(DeviceOps: DeviceArray + C-ABI kernels + CUDA tensors) and, for the world_size-2 gloo tests,
on host arrays (tests/host_ops.py: the oracle + CPU tensors).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist


def pack_bits(torch_mod, valid):
    """bool[n] (n % 8 == 0) -> LSB-first bitmap bytes (torch plumbing for validity columns)."""
    w = torch_mod.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch_mod.uint8, device=valid.device)
    return (valid.view(-1, 8).to(torch_mod.uint8) * w).sum(dim=1, dtype=torch_mod.uint8)


# ------------------------------------------------------------------------------------------------
# collectives on raw torch tensors
# ------------------------------------------------------------------------------------------------
def all_to_all_v(send: torch.Tensor, send_counts: Sequence[int], group=None) -> Tuple[torch.Tensor, List[int]]:
    """Variable-size all-to-all of a 1-D tensor partitioned by destination rank."""
    world = dist.get_world_size(group)
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    return recv, recv_counts


def all_gather_v(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.empty(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    padded = torch.zeros(m, dtype=t.dtype, device=t.device)
    padded[: t.numel()] = t
    outs = [torch.empty(m, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(outs, padded, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


# ------------------------------------------------------------------------------------------------
# the two distributed operators (backend agnostic)
# ------------------------------------------------------------------------------------------------
def group_by_sum_count(keys, values, ops, group=None):
    """Distributed `group_by(key).aggregate(sum, count)` over row-range shards.
    Returns this rank's owned groups: (keys, sums, counts) -- disjoint across ranks, union = all."""
    world = dist.get_world_size(group)
    k, s, c = ops.local_group_by(keys, values)              # partial groups of this shard
    dest = ops.hash_partition(k, world)                     # owner of every partial group
    order = ops.stable_sort_indices(dest)
    k, s, c, dest_sorted = ops.take(k, order), ops.take(s, order), ops.take(c, order), ops.take(dest, order)
    send_counts = ops.histogram(dest_sorted, world)
    k_t, k_null_t = ops.key_tensors(k)                      # raw keys + per-key null flag (uint8)
    s_t, c_t = ops.sum_tensor(s), ops.count_tensor(c)       # null sums travel as 0 (count says it all)
    rk, _ = all_to_all_v(k_t, send_counts, group)
    rn, _ = all_to_all_v(k_null_t, send_counts, group)
    rs, _ = all_to_all_v(s_t, send_counts, group)
    rc, _ = all_to_all_v(c_t, send_counts, group)
    return ops.merge_partials(rk, rn, rs, rc, keys_type=ops.type_of(k), sum_type=ops.type_of(s))


def sort_indices(values, ops, group=None, samples_per_rank: int = 4096):
    """Distributed stable `sort_indices(values, ascending, nulls at end)` over row-range shards.
    Returns (sorted_global_indices_segment, null_global_indices_segment) for this rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = ops.length(values)
    lens = all_gather_v(ops.scalar_tensor(n_local), group)
    row0 = sum(int(x.item()) for x in lens[:rank])
    # splitters from an all-gathered sample of the valid rows
    sample = ops.sample_valid(values, samples_per_rank)
    gathered = torch.cat(all_gather_v(ops.values_tensor(sample), group))
    splitters = ops.pick_splitters(gathered, world, ops.type_of(values))
    dest = ops.range_partition(values, splitters)           # nulls get id world (= P-1 splitters + 1)
    order = ops.stable_sort_indices(dest)
    dest_sorted = ops.take(dest, order)
    counts = ops.histogram(dest_sorted, world + 1)
    n_valid = sum(counts[:world])
    gidx = ops.add_offset(order, row0)                      # global row numbers, partition order
    valid_part = ops.slice(ops.take(values, order), 0, n_valid)
    rk, _ = all_to_all_v(ops.values_tensor(valid_part), counts[:world], group)
    ri, _ = all_to_all_v(ops.index_tensor(ops.slice(gidx, 0, n_valid)), counts[:world], group)
    null_idx = ops.index_tensor(ops.slice(gidx, n_valid, n_local - n_valid))
    # received rows are grouped by source rank (ascending) and ascending global row inside each
    # group, so a STABLE local sort keeps ties in global row order
    local = ops.stable_sort_indices(ops.from_values_tensor(rk, ops.type_of(values)))
    return ops.take_tensor(ri, local), null_idx


# ------------------------------------------------------------------------------------------------
# device backend
# ------------------------------------------------------------------------------------------------
class _CudaView:
    """__cuda_array_interface__ adapter: lets torch view a pool buffer without a copy."""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {pa.int8(): "|i1", pa.uint8(): "|u1", pa.int16(): "<i2", pa.uint16(): "<u2", pa.int32(): "<i4", pa.uint32(): "<u4",
            pa.int64(): "<i8", pa.uint64(): "<u8", pa.float32(): "<f4", pa.float64(): "<f8"}
_TORCH = {pa.int8(): torch.int8, pa.uint8(): torch.uint8, pa.int16(): torch.int16, pa.int32(): torch.int32, pa.int64(): torch.int64,
          pa.float32(): torch.float32, pa.float64(): torch.float64, pa.uint16(): torch.int16, pa.uint32(): torch.int32,
          pa.uint64(): torch.int64}


class DeviceOps:
    """ops interface over DeviceArray + the C-ABI kernels (everything stays in HBM)."""

    def __init__(self, ctx=None):
        from . import compute as bc
        from .device import Context, DeviceArray
        self.bc, self.DeviceArray = bc, DeviceArray
        self.ctx = ctx or Context.get(torch.cuda.current_device())

    # -- array <-> tensor (zero copy) --
    def _tensor(self, arr, t=None):
        t = t or arr.type
        n = arr.length
        if n == 0:
            return torch.empty(0, dtype=_TORCH[t], device="cuda")
        width = t.bit_width // 8
        view = _CudaView(arr.buffers[1].ptr + arr.offset * width, n, _TYPESTR[t], arr)
        out = torch.as_tensor(view, device="cuda")
        return out.view(_TORCH[t]) if out.dtype != _TORCH[t] else out

    def _array(self, tensor, t, validity=None, null_count=0):
        tensor = tensor.contiguous()
        return self.DeviceArray.from_pointers(self.ctx, t, tensor.numel(), tensor.data_ptr() if tensor.numel() else 0,
                                              validity_ptr=validity.data_ptr() if validity is not None else 0,
                                              null_count=null_count, keepalive=(tensor, validity))

    def type_of(self, arr):
        return arr.type

    def length(self, arr):
        return arr.length

    def scalar_tensor(self, v):
        return torch.tensor([v], dtype=torch.int64, device="cuda")

    # -- kernels --
    def local_group_by(self, keys, values):
        g = self.bc.GroupBySumCount(keys.type, values.type, ctx=self.ctx)
        g.consume(keys, values)
        return g.finalize()

    def _partition_call(self, fn_name, *cargs):
        from . import _cabi as cabi
        from .device import check
        out = cabi.B2Array()
        check(getattr(self.ctx.lib, fn_name)(self.ctx.handle, *cargs, C.byref(out), self.ctx.stream))
        return self.DeviceArray._from_c(self.ctx, out, pa.uint32())

    def hash_partition(self, keys, n_parts):
        ck = keys._c()
        return self._partition_call("b2_hash_partition", C.byref(ck), int(n_parts))

    def range_partition(self, values, splitters):
        cv, cs = values._c(), splitters._c()
        return self._partition_call("b2_range_partition", C.byref(cv), C.byref(cs), 0)

    def stable_sort_indices(self, arr):
        return self.bc.array_sort_indices(arr)

    def take(self, arr, idx):
        return self.bc.take(arr, idx)

    def slice(self, arr, off, length):
        return arr.slice(off, length)

    def histogram(self, sorted_ids, n_bins):
        t = self._tensor(sorted_ids, pa.uint32()).to(torch.int64)
        return [int(x) for x in torch.bincount(t, minlength=n_bins)[:n_bins].tolist()]

    def key_tensors(self, k):
        raw = self._tensor(k)
        if k.null_count == 0 or k.buffers[0] is None:
            return raw, torch.zeros(k.length, dtype=torch.uint8, device="cuda")
        valid = self._tensor(self._valid_as_uint8(k))
        return raw, (1 - valid).to(torch.uint8)

    def _valid_as_uint8(self, arr):
        # validity bitmap -> uint8 0/1 column: compare(equal(arr, arr)) is all-true where valid and null
        # elsewhere; take its validity by counting through a filter-free path: unpack with torch
        bits = _CudaView(arr.buffers[0].ptr, (arr.offset + arr.length + 7) // 8, "|u1", arr)
        b = torch.as_tensor(bits, device="cuda")
        shifts = torch.arange(8, device="cuda", dtype=torch.uint8)
        un = ((b.unsqueeze(1) >> shifts) & 1).reshape(-1)[arr.offset: arr.offset + arr.length].contiguous()
        return self._array(un, pa.uint8())

    def sum_tensor(self, s):
        t = self._tensor(s)
        if s.null_count and s.buffers[0] is not None:
            t = t * self._tensor(self._valid_as_uint8(s)).to(t.dtype)
        return t

    def count_tensor(self, c):
        return self._tensor(c)

    def values_tensor(self, arr):
        return self._tensor(arr)

    def index_tensor(self, arr):
        return self._tensor(arr, pa.uint64()) if arr.type == pa.uint64() else self._tensor(arr)

    def from_values_tensor(self, t, typ):
        return self._array(t, typ)

    def take_tensor(self, t, idx_arr):
        return t[self._tensor(idx_arr, pa.uint64())]

    def add_offset(self, idx_arr, off):
        t = self._tensor(idx_arr, pa.uint64()) + off
        return self._array(t, pa.uint64())

    def sample_valid(self, values, k):
        n = values.length
        if n == 0:
            return values
        step = max(1, n // k)
        idx = torch.arange(0, n, step, dtype=torch.int64, device="cuda")
        s = self.bc.take(values, self._array(idx, pa.int64()))
        if s.null_count:  # drop sampled nulls: filter with the sample's own validity
            keep = self._array((self._tensor(self._valid_as_uint8(s)) != 0).to(torch.uint8), pa.uint8())
            s = self.bc.filter(s, self.bc.not_equal(keep, 0))
        return s

    def pick_splitters(self, gathered, world, typ):
        g, _ = torch.sort(gathered)
        if g.numel() == 0 or world == 1:
            return self._array(g[:0], typ)
        pos = (torch.arange(1, world, device=g.device) * g.numel()) // world
        return self._array(g[pos.clamp(max=g.numel() - 1)], typ)

    def merge_partials(self, rk, rn, rs, rc, keys_type, sum_type):
        n = rk.numel()
        validity = None
        nulls = int(rn.sum().item()) if n else 0
        if nulls:
            m8 = (n + 7) // 8 * 8
            v = torch.zeros(m8, dtype=torch.bool, device="cuda")
            v[:n] = rn == 0
            validity = pack_bits(torch, v)
        keys = self._array(rk, keys_type, validity, nulls)
        g = self.bc.Grouper([keys_type], self.ctx)
        ids = g.consume(keys)
        sums = self.bc.HashAggregator("hash_sum", sum_type, ctx=self.ctx)
        cnts = self.bc.HashAggregator("hash_sum", pa.int64(), ctx=self.ctx)
        sums.resize(g.num_groups)
        cnts.resize(g.num_groups)
        sums.consume(self._array(rs, sum_type), ids)
        cnts.consume(self._array(rc, pa.int64()), ids)
        total = cnts.finalize()
        s = sums.finalize()
        # a group whose merged count is 0 has a null sum (min_count = 1): re-mask from the counts
        zero = self.bc.equal(total, 0)
        nz = self.bc.filter_output_size(zero)
        if nz:
            cnt_t = self._tensor(total)
            m8 = (cnt_t.numel() + 7) // 8 * 8
            v = torch.zeros(m8, dtype=torch.bool, device="cuda")
            v[: cnt_t.numel()] = cnt_t != 0
            s = self._array(self._tensor(s).clone(), sum_type, pack_bits(torch, v), nz)
        return g.get_uniques()[0], s, total
