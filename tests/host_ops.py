"""Host stand-in for arrow_b200.distributed.DeviceOps, used only by the world_size-2 gloo tests:
the same `ops` interface over pyarrow host arrays, the oracle and CPU tensors, so the exchange
logic (partition, splits, all-to-all, merge, ordering) is exercised without a GPU."""
import numpy as np
import pyarrow as pa
import torch

from oracle import arrow_oracle as ora

_MIX = (0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53)


def hash64(k: np.ndarray) -> np.ndarray:
    """murmur3 fmix64 -- must equal hash64() in arrow_b200/csrc/hash_table.cuh"""
    k = k.astype(np.uint64)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(_MIX[0])
        k ^= k >> np.uint64(33)
        k *= np.uint64(_MIX[1])
        k ^= k >> np.uint64(33)
    return k


class HostOps:
    def type_of(self, arr):
        return arr.type

    def length(self, arr):
        return len(arr)

    def scalar_tensor(self, v):
        return torch.tensor([v], dtype=torch.int64)

    def local_group_by(self, keys, values):
        uniq, (s, c) = ora.group_by([keys], [("hash_sum", values, None), ("hash_count", values, None)])
        return uniq[0], s, c

    def hash_partition(self, keys, n_parts):
        v, valid = ora.values(keys), ora.validity(keys)
        raw = v.view(np.uint64) if v.dtype.itemsize == 8 else v.astype(np.uint64)
        ids = (hash64(raw) % np.uint64(n_parts)).astype(np.uint32)
        return pa.array(np.where(valid, ids, 0).astype(np.uint32))

    def range_partition(self, values, splitters):
        v, valid = ora.values(values), ora.validity(values)
        sp = ora.values(splitters)
        ids = np.searchsorted(sp, v, side="right").astype(np.uint32)
        return pa.array(np.where(valid, ids, len(sp) + 1).astype(np.uint32))

    def stable_sort_indices(self, arr):
        return ora.sort_indices(arr)

    def take(self, arr, idx):
        return ora.take(arr, idx)

    def slice(self, arr, off, length):
        return arr.slice(off, length)

    def histogram(self, sorted_ids, n_bins):
        return [int(x) for x in np.bincount(ora.values(sorted_ids), minlength=n_bins)[:n_bins]]

    def key_tensors(self, k):
        return torch.from_numpy(ora.values(k).copy()), torch.from_numpy((~ora.validity(k)).astype(np.uint8))

    def sum_tensor(self, s):
        return torch.from_numpy(np.where(ora.validity(s), ora.values(s), 0).copy())

    def count_tensor(self, c):
        return torch.from_numpy(ora.values(c).copy())

    def values_tensor(self, arr):
        return torch.from_numpy(ora.values(arr).copy())

    def index_tensor(self, arr):
        return torch.from_numpy(ora.values(arr).astype(np.int64))

    def from_values_tensor(self, t, typ):
        return pa.array(t.numpy(), typ)

    def take_tensor(self, t, idx_arr):
        return t[torch.from_numpy(ora.values(idx_arr).astype(np.int64))]

    def add_offset(self, idx_arr, off):
        return pa.array(ora.values(idx_arr).astype(np.uint64) + np.uint64(off), pa.uint64())

    def sample_valid(self, values, k):
        n = len(values)
        if n == 0:
            return values
        s = values.take(pa.array(np.arange(0, n, max(1, n // k))))
        return s.drop_null()

    def pick_splitters(self, gathered, world, typ):
        g, _ = torch.sort(gathered)
        if g.numel() == 0 or world == 1:
            return pa.array(g[:0].numpy(), typ)
        pos = (torch.arange(1, world) * g.numel()) // world
        return pa.array(g[pos.clamp(max=g.numel() - 1)].numpy(), typ)

    def merge_partials(self, rk, rn, rs, rc, keys_type, sum_type):
        keys = pa.array(rk.numpy(), keys_type, mask=rn.numpy().astype(bool))
        uniq, (s, c) = ora.group_by([keys], [("hash_sum", pa.array(rs.numpy(), sum_type), None),
                                              ("hash_sum", pa.array(rc.numpy(), pa.int64()), None)])
        cv = ora.values(c)
        s = ora.make_array(sum_type, ora.values(s), cv != 0)
        return uniq[0], s, c
