// grouper.cu -- key columns -> dense uint32 group ids, on device.
//
// Replaces arrow::compute::Grouper (cpp/src/arrow/compute/row/grouper.h:104-196) as
// implemented by GrouperFastImpl (row/grouper.cc:555-963: row encode -> Hashing32 ->
// SwissTable early_filter/find -> map_new_keys) for fixed-width key columns:
//   Make / Consume / Lookup / GetUniques / num_groups / Reset.
// Semantics kept: one id per row, ids are dense [0, num_groups), a null key is its own
// group, keys group by BYTES (so -0.0 != +0.0 and NaNs group by payload, row/grouper.cc
// encodes raw bytes), uniques only ever grow by appending (prefix-stable), Lookup never
// inserts and yields null for unknown keys.  Stronger than the reference: ids are
// assigned in first-occurrence row order, deterministically (the reference only
// guarantees a bijection of that order).
//
// B200 design (per Consume call, all stream-ordered):
//   insert : every row encodes its key (<= 64 bits incl. in-band null flags) and finds or claims
//            its 16-byte slot {key, id, first_row} (linear probing, one 16-byte load per probe,
//            one 64-bit atomicCAS to claim).  A group that already has an id (seen in an earlier
//            batch) yields the row's id right away; a group that is new in this batch lowers the
//            slot's first_row to the row number (atomicMin, skipped when a smaller row got there
//            first -- the common case since rows are visited in ascending order).
//   flag   : bit i = row i is the first occurrence of a new group.  Built from the table (one
//            streaming pass over the slots) when the table is not larger than the batch, else
//            from the unresolved rows.
//   rank   : the Filter count+scan pass turns the flag bitmap into per-tile offsets, so
//            new-group ids = num_groups + rank (first-occurrence order, no atomics).
//   assign : first-occurrence rows publish id and append the encoded key to `uniques`.
//   gather : out_ids[i] = id[slot[i]] for the rows the insert pass left unresolved.
// Lookup is one pass (probe + id read + validity ballot).  The table grows 4x (rehash from
// `uniques`) when the probe limit is hit -- the insert pass stops at the first overflow -- or
// when the load exceeds 1/2.  Every random access is one 32-byte sector: measured 25.5 ms
// (insert) + 22.1 ms (gather) per 1B rows at 10M groups = the 42 G accesses/s DRAM ceiling.
//
// Direct-addressed mode (one integer key column whose values span <= 2^24): the hash table is
// replaced by two uint32 arrays indexed by key - window_lo, {id} and {first_row}, each small enough
// to stay L2-resident, so the two random accesses per row never reach DRAM.  The window is exact
// (a min/max pass over the batch precedes every Consume), grows by re-seating the uniques, and the
// grouper falls back to the hash table -- rebuilt from the same uniques -- the first time a batch
// does not fit.  Ids, uniques and Lookup results are identical in both modes.
#include <cstdlib>

#include "hash_table.cuh"
#include "selection.cuh"

namespace b2 {

constexpr uint32_t kNoId = 0xffffffffu;
constexpr int kMaxKeys = 8;

struct KeyLayout {
  int n_keys;
  int width[kMaxKeys];    // bytes
  int bit_off[kMaxKeys];  // position of column j inside the encoded key
  int null_bit[kMaxKeys]; // in-band null flag position (multi-column), -1 if single column
};

struct KeyColumns {
  const void* data[kMaxKeys];  // advanced by offset
  BitmapReader valid[kMaxKeys];
};

__device__ __forceinline__ uint64_t encode_row(const KeyLayout& L, const KeyColumns& c, int64_t i, bool* is_null) {
  uint64_t enc = 0;
  *is_null = false;
  if (L.n_keys == 1) {
    if (!c.valid[0].bit(i)) {
      *is_null = true;
      return 0;
    }
    return load_key_bits(c.data[0], L.width[0], i);
  }
#pragma unroll 1
  for (int j = 0; j < L.n_keys; ++j) {
    if (c.valid[j].bit(i)) enc |= load_key_bits(c.data[j], L.width[j], i) << L.bit_off[j];
    else enc |= 1ull << L.null_bit[j];
  }
  return enc;
}

// 16-byte slots {key, id, first_row}: the probe, the id read and the first-occurrence atomicMin of
// a row all land in ONE 32-byte sector (three separate arrays cost three random DRAM accesses).
struct GrouperTable {
  unsigned long long* slots;  // [(cap + 2) * 2]: word 0 = key, word 1 = {id, first_row}
  uint64_t mask;
  __host__ __device__ uint32_t* id_ptr(uint64_t slot) const { return reinterpret_cast<uint32_t*>(slots + slot * 2 + 1); }
  __host__ __device__ uint32_t* first_row_ptr(uint64_t slot) const { return id_ptr(slot) + 1; }
};

// direct-addressed mode (see the header comment)
constexpr uint64_t kDirectMaxRange = 1ull << 24;

struct DirectTable {
  uint32_t* id;         // [cap + 1]; entry `cap` is the null group
  uint32_t* first_row;  // [cap + 1]
  uint64_t lo;          // order-preserving encoding of entry 0
  uint64_t cap;
  uint64_t flip;        // sign bit of a signed key type (makes the unsigned order the value order)
  __host__ __device__ uint64_t index(uint64_t enc, bool is_null) const { return is_null ? cap : (enc ^ flip) - lo; }
};

__global__ void __launch_bounds__(kBlock) grouper_init_kernel(GrouperTable t) {
  for (uint64_t i = blockIdx.x * (uint64_t)kBlock + threadIdx.x; i < t.mask + 3; i += (uint64_t)gridDim.x * kBlock) {
    t.slots[i * 2] = kEmptyKey;
    t.slots[i * 2 + 1] = ~0ull;  // id = kNoId, first_row = 0xffffffff
  }
}

constexpr uint32_t kResolved = 0xfffffffeu;  // row_slot marker: the id was already written by the probe

// Consume probe: find or claim the slot of every row.  A row of a group that already has an
// id (seen in an earlier batch) writes its id right away; a row of a group that is new in this
// batch remembers its slot for the gather pass and lowers the slot's first_row to its own row
// number (the atomicMin is skipped when a smaller row got there first, which is the common case
// because rows are visited in ascending order).  Stops early once the table overflowed.
__global__ void __launch_bounds__(kBlock) grouper_insert_kernel(KeyLayout L, KeyColumns c, int64_t n,
                                                                GrouperTable t, uint32_t* row_slot,
                                                                uint32_t* out_ids, int64_t* overflow) {
  const volatile int64_t* ovf = overflow;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    if (*ovf) return;
    bool is_null;
    uint64_t enc = encode_row(L, c, i, &is_null);
    int64_t slot = -1;
    unsigned long long w1 = ~0ull;
    if (is_null || enc == kEmptyKey) {
      bool inserted;
      slot = table_find_or_insert(t.slots, t.mask, 2, enc, is_null, &inserted);
      w1 = *reinterpret_cast<volatile unsigned long long*>(t.slots + slot * 2 + 1);
    } else {
      // one 16-byte load per probe brings the key and {id, first_row} together
      uint64_t sl = hash64(enc) & t.mask;
      for (int probe = 0; probe < kMaxProbe; ++probe) {
        unsigned long long* p = t.slots + sl * 2;
        unsigned long long k, v;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(k), "=l"(v) : "l"(p) : "memory");
        if (k == kEmptyKey) {
          k = atomicCAS(p, (unsigned long long)kEmptyKey, (unsigned long long)enc);
          if (k == kEmptyKey) {  // claimed: a fresh slot has no id and no first row yet
            slot = static_cast<int64_t>(sl);
            break;
          }
          if (k == enc) v = *reinterpret_cast<volatile unsigned long long*>(p + 1);
        }
        if (k == enc) {
          slot = static_cast<int64_t>(sl);
          w1 = v;
          break;
        }
        sl = (sl + 1) & t.mask;
      }
    }
    if (slot < 0) {
      *overflow = 1;
      return;
    }
    const uint32_t id = static_cast<uint32_t>(w1), first_row = static_cast<uint32_t>(w1 >> 32);
    if (id != kNoId) {
      out_ids[i] = id;
      row_slot[i] = kResolved;
    } else {
      if (first_row > static_cast<uint32_t>(i)) atomicMin(t.first_row_ptr(slot), static_cast<uint32_t>(i));
      row_slot[i] = static_cast<uint32_t>(slot);
    }
  }
}

// Lookup: never inserts; unknown keys become null (validity via ballot)
__global__ void __launch_bounds__(kBlock) grouper_lookup_kernel(KeyLayout L, KeyColumns c, int64_t n,
                                                                GrouperTable t, uint32_t* out,
                                                                uint32_t* out_validity, int64_t* valid_count) {
  int64_t nw = (n + 31) >> 5;
  int64_t local = 0;
  for (int64_t w = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5; w < nw; w += ((int64_t)gridDim.x * kBlock) >> 5) {
    int64_t i = (w << 5) + lane_id();
    bool ok = false;
    if (i < n) {
      bool is_null;
      uint64_t enc = encode_row(L, c, i, &is_null);
      int64_t slot = table_find(t.slots, t.mask, 2, enc, is_null);
      uint32_t id = slot < 0 ? kNoId : *t.id_ptr(slot);
      ok = id != kNoId;
      out[i] = ok ? id : 0u;
    }
    unsigned word = __ballot_sync(0xffffffffu, ok);
    if (lane_id() == 0) {
      out_validity[w] = word;
      local += __popc(word);
    }
  }
  int64_t s = block_sum<kBlock>(local);
  if (threadIdx.x == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(valid_count), (unsigned long long)s);
}

// bit i = row i is the first occurrence of a new group.  Two ways to build the bitmap:
//   by rows  : every unresolved row re-reads its slot (one random access per row);
//   by slots : one streaming pass over the table sets bit first_row of every slot that
//              has no id yet -- cheaper whenever the table is not much larger than the batch.
__global__ void __launch_bounds__(kBlock) grouper_flag_rows_kernel(int64_t n, GrouperTable t,
                                                                   const uint32_t* row_slot, uint32_t* flags) {
  int64_t nw = (n + 31) >> 5;
  for (int64_t w = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5; w < nw; w += ((int64_t)gridDim.x * kBlock) >> 5) {
    int64_t i = (w << 5) + lane_id();
    bool f = false;
    if (i < n) {
      uint32_t s = row_slot[i];
      if (s != kResolved) {
        const unsigned long long w1 = t.slots[(uint64_t)s * 2 + 1];
        f = static_cast<uint32_t>(w1) == kNoId && static_cast<uint32_t>(w1 >> 32) == static_cast<uint32_t>(i);
      }
    }
    unsigned word = __ballot_sync(0xffffffffu, f);
    if (lane_id() == 0) flags[w] = word;
  }
}

__global__ void __launch_bounds__(kBlock) grouper_flag_slots_kernel(GrouperTable t, uint32_t* flags) {
  for (uint64_t s = blockIdx.x * (uint64_t)kBlock + threadIdx.x; s < t.mask + 3; s += (uint64_t)gridDim.x * kBlock) {
    const unsigned long long w1 = __ldcs(t.slots + s * 2 + 1);
    const uint32_t id = static_cast<uint32_t>(w1), first_row = static_cast<uint32_t>(w1 >> 32);
    if (id == kNoId && first_row != 0xffffffffu) atomicOr(flags + (first_row >> 5), 1u << (first_row & 31));
  }
}

template <bool DIRECT>
__global__ void __launch_bounds__(kBlock) grouper_assign_kernel(KeyLayout L, KeyColumns c, int64_t n,
                                                                GrouperTable t, DirectTable d, const uint32_t* row_slot,
                                                                BitmapReader flags, const int64_t* tile_offsets,
                                                                uint32_t base_id, uint64_t* uniq_keys,
                                                                uint8_t* uniq_null) {
  __shared__ uint64_t s_sel[kTileWords];
  __shared__ uint32_t s_prefix[kTileWords];
  const int64_t tile = blockIdx.x;
  const int64_t row0 = tile * kTileRows;
  const unsigned lane = lane_id();
  if (threadIdx.x < 32) {
    int64_t w0 = tile * kTileWords + 2 * lane;
    uint64_t s0 = flags.word(w0), s1 = flags.word(w0 + 1);
    int c0 = __popcll(s0), c1 = __popcll(s1);
    int incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int excl = incl - c0 - c1;
    s_sel[2 * lane] = s0;
    s_sel[2 * lane + 1] = s1;
    s_prefix[2 * lane] = excl;
    s_prefix[2 * lane + 1] = excl + c0;
  }
  __syncthreads();
  const int64_t vbase = tile_offsets[tile];
  for (int p = 0; p < kTileRows / kBlock; ++p) {
    const int r = p * kBlock + threadIdx.x;
    const int64_t row = row0 + r;
    if (row >= n) break;
    const uint64_t selw = s_sel[r >> 6];
    if (!((selw >> (r & 63)) & 1)) continue;
    const unsigned rank = s_prefix[r >> 6] + __popcll(selw & ((1ull << (r & 63)) - 1ull));
    const uint32_t id = base_id + static_cast<uint32_t>(vbase + rank);
    bool is_null;
    uint64_t enc = encode_row(L, c, row, &is_null);
    if (DIRECT) d.id[d.index(enc, is_null)] = id;
    else *t.id_ptr(row_slot[row]) = id;
    uniq_keys[id] = enc;
    uniq_null[id] = is_null ? 1 : 0;
  }
}

// out[i] = id of row i for the rows the probe left unresolved (groups new in this batch)
__global__ void __launch_bounds__(kBlock) grouper_gather_kernel(int64_t n, GrouperTable t,
                                                                const uint32_t* __restrict__ row_slot,
                                                                uint32_t* __restrict__ out) {
  constexpr int kPer = 4;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  for (int64_t i0 = blockIdx.x * (int64_t)kBlock + threadIdx.x; i0 < n; i0 += stride * kPer) {
    uint32_t s[kPer], id[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int64_t i = i0 + k * stride;
      s[k] = i < n ? __ldcs(row_slot + i) : kResolved;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (s[k] != kResolved) id[k] = *t.id_ptr(s[k]);
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (s[k] != kResolved) out[i0 + k * stride] = id[k];
  }
}

__global__ void __launch_bounds__(kBlock) grouper_rehash_kernel(GrouperTable t, const uint64_t* uniq_keys,
                                                                const uint8_t* uniq_null, uint32_t n_groups,
                                                                int64_t* overflow) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < n_groups; g += gridDim.x * kBlock) {
    bool inserted;
    int64_t slot = table_find_or_insert(t.slots, t.mask, 2, uniq_keys[g], uniq_null[g] != 0, &inserted);
    if (slot < 0) {
      *overflow = 1;
      continue;
    }
    *t.id_ptr(slot) = g;
  }
}

// uniques -> one key column
__global__ void __launch_bounds__(kBlock) grouper_decode_kernel(KeyLayout L, int col, const uint64_t* uniq_keys,
                                                                const uint8_t* uniq_null, uint32_t n_groups,
                                                                void* out, uint32_t* out_validity,
                                                                int64_t* valid_count) {
  uint32_t nw = (n_groups + 31) >> 5;
  int64_t local = 0;
  for (uint32_t w = (blockIdx.x * kBlock + threadIdx.x) >> 5; w < nw; w += (gridDim.x * kBlock) >> 5) {
    uint32_t g = (w << 5) + lane_id();
    bool valid = false;
    if (g < n_groups) {
      uint64_t enc = uniq_keys[g];
      uint64_t v;
      if (L.n_keys == 1) {
        valid = uniq_null[g] == 0;
        v = enc;
      } else {
        valid = !((enc >> L.null_bit[col]) & 1);
        v = enc >> L.bit_off[col];
      }
      if (!valid) v = 0;
      switch (L.width[col]) {
        case 1: static_cast<uint8_t*>(out)[g] = static_cast<uint8_t>(v); break;
        case 2: static_cast<uint16_t*>(out)[g] = static_cast<uint16_t>(v); break;
        case 4: static_cast<uint32_t*>(out)[g] = static_cast<uint32_t>(v); break;
        default: static_cast<uint64_t*>(out)[g] = v; break;
      }
    }
    unsigned word = __ballot_sync(0xffffffffu, valid);
    if (lane_id() == 0) {
      out_validity[w] = word;
      local += __popc(word);
    }
  }
  int64_t s = block_sum<kBlock>(local);
  if (threadIdx.x == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(valid_count), (unsigned long long)s);
}

// ---------------------------------------------------------------------------
// direct-addressed mode
// ---------------------------------------------------------------------------
// mm[0] = min, mm[1] = max of the order-preserving encoding over the valid rows
__global__ void __launch_bounds__(kBlock) grouper_minmax_kernel(const void* data, int width, BitmapReader valid,
                                                                int64_t n, uint64_t flip, unsigned long long* mm) {
  uint64_t lo = ~0ull, hi = 0;
  constexpr int kPer = 4;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  for (int64_t i0 = blockIdx.x * (int64_t)kBlock + threadIdx.x; i0 < n; i0 += stride * kPer) {
    uint64_t v[kPer];
    bool ok[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int64_t i = i0 + k * stride;
      ok[k] = i < n && valid.bit(i);
      v[k] = ok[k] ? load_key_bits(data, width, i) ^ flip : 0;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (!ok[k]) continue;
      lo = v[k] < lo ? v[k] : lo;
      hi = v[k] > hi ? v[k] : hi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if (lane_id() == 0 && lo <= hi) {
    atomicMin(mm, (unsigned long long)lo);
    atomicMax(mm + 1, (unsigned long long)hi);
  }
}

// Pass A: rows of groups that have no id yet lower first_row (guarded, so almost always a plain L2
// read); unless the grouper is FRESH, rows of known groups get their id here.
template <bool FRESH>
__global__ void __launch_bounds__(kBlock) grouper_direct_first_kernel(const void* data, int width, BitmapReader valid,
                                                                      int64_t n, DirectTable t, uint32_t* out_ids) {
  constexpr int kPer = 4;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  for (int64_t i0 = blockIdx.x * (int64_t)kBlock + threadIdx.x; i0 < n; i0 += stride * kPer) {
    uint64_t idx[kPer];
    uint32_t id[kPer], fr[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int64_t i = i0 + k * stride;
      idx[k] = ~0ull;
      if (i < n) idx[k] = t.index(load_key_bits(data, width, i), !valid.bit(i));
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      id[k] = kNoId;
      if (!FRESH && idx[k] != ~0ull) id[k] = t.id[idx[k]];
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      fr[k] = 0;
      if (idx[k] != ~0ull && id[k] == kNoId) fr[k] = *reinterpret_cast<volatile uint32_t*>(t.first_row + idx[k]);
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int64_t i = i0 + k * stride;
      if (idx[k] == ~0ull) continue;
      if (id[k] == kNoId && fr[k] > static_cast<uint32_t>(i)) atomicMin(t.first_row + idx[k], static_cast<uint32_t>(i));
      if (!FRESH) out_ids[i] = id[k];
    }
  }
}

__global__ void __launch_bounds__(kBlock) grouper_direct_flag_kernel(DirectTable t, uint32_t* flags) {
  for (uint64_t e = blockIdx.x * (uint64_t)kBlock + threadIdx.x; e <= t.cap; e += (uint64_t)gridDim.x * kBlock) {
    const uint32_t first_row = __ldcs(t.first_row + e);
    if (first_row != 0xffffffffu && t.id[e] == kNoId) atomicOr(flags + (first_row >> 5), 1u << (first_row & 31));
  }
}

__global__ void __launch_bounds__(kBlock) grouper_direct_gather_kernel(const void* data, int width, BitmapReader valid,
                                                                       int64_t n, DirectTable t,
                                                                       uint32_t* __restrict__ out) {
  constexpr int kPer = 4;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  for (int64_t i0 = blockIdx.x * (int64_t)kBlock + threadIdx.x; i0 < n; i0 += stride * kPer) {
    uint64_t idx[kPer];
    uint32_t id[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int64_t i = i0 + k * stride;
      idx[k] = ~0ull;
      if (i < n) idx[k] = t.index(load_key_bits(data, width, i), !valid.bit(i));
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (idx[k] != ~0ull) id[k] = t.id[idx[k]];
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (idx[k] != ~0ull) __stcs(out + i0 + k * stride, id[k]);
  }
}

__global__ void __launch_bounds__(kBlock) grouper_direct_reseat_kernel(DirectTable t, const uint64_t* uniq_keys,
                                                                       const uint8_t* uniq_null, uint32_t n_groups) {
  for (uint32_t g = blockIdx.x * kBlock + threadIdx.x; g < n_groups; g += gridDim.x * kBlock)
    t.id[t.index(uniq_keys[g], uniq_null[g] != 0)] = g;
}

__global__ void __launch_bounds__(kBlock) grouper_direct_lookup_kernel(const void* data, int width, BitmapReader valid,
                                                                       int64_t n, DirectTable t, uint32_t* out,
                                                                       uint32_t* out_validity, int64_t* valid_count) {
  int64_t nw = (n + 31) >> 5;
  int64_t local = 0;
  for (int64_t w = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5; w < nw; w += ((int64_t)gridDim.x * kBlock) >> 5) {
    int64_t i = (w << 5) + lane_id();
    bool ok = false;
    if (i < n) {
      uint32_t id = kNoId;
      if (!valid.bit(i)) {
        id = t.id[t.cap];
      } else {
        const uint64_t ord = load_key_bits(data, width, i) ^ t.flip;
        if (ord >= t.lo && ord - t.lo < t.cap) id = t.id[ord - t.lo];
      }
      ok = id != kNoId;
      out[i] = ok ? id : 0u;
    }
    unsigned word = __ballot_sync(0xffffffffu, ok);
    if (lane_id() == 0) {
      out_validity[w] = word;
      local += __popc(word);
    }
  }
  int64_t s = block_sum<kBlock>(local);
  if (threadIdx.x == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(valid_count), (unsigned long long)s);
}

// utf8 / binary keys and keys wider than 64 bits: grouper_wide.cu (parts reduced to ids by this file's kernel, then folded)
struct WideGrouper;
int wide_create(B2Context* ctx, const int32_t* key_types, int n_keys, WideGrouper** out);
void wide_destroy(WideGrouper* g);
int wide_reset(WideGrouper* g);
uint32_t wide_num_groups(const WideGrouper* g);
int wide_run(WideGrouper* g, const B2Array* keys, B2Array* out_ids, bool insert, cudaStream_t s);
int wide_uniques(WideGrouper* g, B2Array* out_keys, cudaStream_t s);

}  // namespace b2

using namespace b2;

struct B2Grouper {
  B2Context* ctx;
  WideGrouper* wide = nullptr;  // set: every call is forwarded
  KeyLayout layout;
  int32_t key_types[kMaxKeys];
  GrouperTable table{};
  uint64_t cap = 0;
  uint32_t num_groups = 0;
  uint64_t* uniq_keys = nullptr;
  uint8_t* uniq_null = nullptr;
  uint64_t uniq_cap = 0;
  // direct-addressed mode
  bool direct_eligible = false;  // one integer key column, not yet known to be too wide
  bool never_direct = false;     // set by callers whose keys are hashes (grouper_wide.cu): skip the min/max probing
  bool direct = false;
  DirectTable dt{};
  uint64_t obs_min = ~0ull, obs_max = 0;  // order-preserving encoding, over every valid key seen
};

static bool grouper_direct_enabled() {
  const char* e = std::getenv("B2_GROUPER_DIRECT");
  return !(e && e[0] == '0');
}

static bool grouper_type_is_integer(int32_t t) { return t >= B2_UINT8 && t <= B2_INT64; }

static uint64_t grouper_sign_flip(int32_t t) {
  switch (t) {
    case B2_INT8: return 1ull << 7;
    case B2_INT16: return 1ull << 15;
    case B2_INT32: return 1ull << 31;
    case B2_INT64: return 1ull << 63;
    default: return 0;
  }
}

static void grouper_free_direct(B2Grouper* g, cudaStream_t s) {
  if (g->dt.id) g->ctx->free(g->dt.id, s);
  if (g->dt.first_row) g->ctx->free(g->dt.first_row, s);
  g->dt = DirectTable{};
  g->direct = false;
}

// (re)seat the direct table on a window that covers [mn, mx] with headroom on both sides
static int grouper_build_direct(B2Grouper* g, uint64_t mn, uint64_t mx, cudaStream_t s) {
  const uint64_t span = mx - mn + 1;
  uint64_t cap = next_pow2(span * 2 < 1024 ? 1024 : span * 2);
  if (cap > kDirectMaxRange) cap = kDirectMaxRange;
  const uint64_t pad = (cap - span) / 2;
  uint64_t lo = mn >= pad ? mn - pad : 0;
  if (lo > ~0ull - (cap - 1)) lo = ~0ull - (cap - 1);
  const uint64_t flip = grouper_sign_flip(g->key_types[0]);
  grouper_free_direct(g, s);
  void *a, *b;
  B2_RETURN_NOT_OK(g->ctx->alloc((cap + 1) * 4, &a, s));
  B2_RETURN_NOT_OK(g->ctx->alloc((cap + 1) * 4, &b, s));
  B2_CUDA(cudaMemsetAsync(a, 0xff, (cap + 1) * 4, s));
  B2_CUDA(cudaMemsetAsync(b, 0xff, (cap + 1) * 4, s));
  g->dt.id = static_cast<uint32_t*>(a);
  g->dt.first_row = static_cast<uint32_t*>(b);
  g->dt.lo = lo;
  g->dt.cap = cap;
  g->dt.flip = flip;
  g->direct = true;
  if (g->num_groups) {
    grouper_direct_reseat_kernel<<<grid_for(g->num_groups, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(
        g->dt, g->uniq_keys, g->uniq_null, g->num_groups);
    B2_LAUNCHED();
  }
  return B2_OK;
}

static void grouper_free_table(B2Grouper* g, cudaStream_t s) {
  if (g->table.slots) g->ctx->free(g->table.slots, s);
  g->table = GrouperTable{};
  g->cap = 0;
}

static int grouper_alloc_table(B2Grouper* g, uint64_t cap, cudaStream_t s) {
  GrouperTable t;
  void* p;
  B2_RETURN_NOT_OK(g->ctx->alloc((cap + 2) * 16, &p, s));
  t.slots = static_cast<unsigned long long*>(p);
  t.mask = cap - 1;
  grouper_init_kernel<<<grid_for((int64_t)cap + 2, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(t);
  B2_LAUNCHED();
  g->table = t;
  g->cap = cap;
  return B2_OK;
}

// (re)build the table at `cap` slots from the uniques
static int grouper_rebuild(B2Grouper* g, uint64_t cap, cudaStream_t s) {
  while (true) {
    grouper_free_table(g, s);
    B2_RETURN_NOT_OK(grouper_alloc_table(g, cap, s));
    if (g->num_groups == 0) return B2_OK;
    ScalarSlot slot(g->ctx);
    B2_RETURN_NOT_OK(slot.zero(s));
    grouper_rehash_kernel<<<grid_for(g->num_groups, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(
        g->table, g->uniq_keys, g->uniq_null, g->num_groups, slot.dev());
    B2_LAUNCHED();
    B2_RETURN_NOT_OK(slot.fetch(s));
    if (!slot.host()[0]) return B2_OK;
    cap *= 4;
  }
}

static int grouper_reserve_uniques(B2Grouper* g, uint64_t need, cudaStream_t s) {
  if (need <= g->uniq_cap) return B2_OK;
  uint64_t cap = next_pow2(need < 1024 ? 1024 : need);
  void *k, *nl;
  B2_RETURN_NOT_OK(g->ctx->alloc(cap * 8, &k, s));
  B2_RETURN_NOT_OK(g->ctx->alloc(cap, &nl, s));
  if (g->num_groups) {
    B2_CUDA(cudaMemcpyAsync(k, g->uniq_keys, (size_t)g->num_groups * 8, cudaMemcpyDeviceToDevice, s));
    B2_CUDA(cudaMemcpyAsync(nl, g->uniq_null, (size_t)g->num_groups, cudaMemcpyDeviceToDevice, s));
  }
  if (g->uniq_keys) g->ctx->free(g->uniq_keys, s);
  if (g->uniq_null) g->ctx->free(g->uniq_null, s);
  g->uniq_keys = static_cast<uint64_t*>(k);
  g->uniq_null = static_cast<uint8_t*>(nl);
  g->uniq_cap = cap;
  return B2_OK;
}

static int grouper_columns(const B2Grouper* g, const B2Array* keys, KeyColumns* c, int64_t* n) {
  if (!keys) return set_error(B2_INVALID, "grouper: null keys");
  *n = keys[0].length;
  for (int j = 0; j < g->layout.n_keys; ++j) {
    if (keys[j].type != g->key_types[j])
      return set_error(B2_INVALID, "expected batch value %d of type id %d but got %d", j, g->key_types[j], keys[j].type);
    if (keys[j].length != *n) return set_error(B2_INVALID, "grouper: key columns differ in length");
    c->data[j] = static_cast<const char*>(keys[j].data) + keys[j].offset * g->layout.width[j];
    c->valid[j] = BitmapReader(keys[j].null_count == 0 ? nullptr : keys[j].validity, keys[j].offset, keys[j].length);
  }
  return B2_OK;
}

// rows (B2_UINT32, ascending = id order) at which the groups NEW in this batch first occur: the set bits of `flags`
static int grouper_new_rows(B2Context* ctx, const void* flags, int64_t n, B2Array* out_rows, cudaStream_t s) {
  B2Array mask{};
  mask.type = B2_BOOL;
  mask.data = flags;
  mask.length = n;
  B2Array rows{};
  B2_RETURN_NOT_OK(b2_filter_indices(ctx, &mask, 0, &rows, s));
  if (rows.type == B2_UINT32) {
    *out_rows = rows;
    return B2_OK;
  }
  B2CastOptions wide{B2_UINT32, 1, 1, 0};  // short batches come back as uint16
  const int st = b2_cast_numeric(ctx, &rows, &wide, out_rows, s);
  if (rows.data) ctx->free(const_cast<void*>(rows.data), s);
  if (rows.validity) ctx->free(const_cast<void*>(rows.validity), s);
  return st;
}

static int grouper_run(B2Grouper* g, const B2Array* keys, B2Array* out_ids, bool insert, cudaStream_t s,
                       B2Array* out_new_rows = nullptr) {
  B2Context* ctx = g->ctx;
  if (out_new_rows) fill_out(out_new_rows, B2_UINT32, 0, 0, nullptr, nullptr);
  KeyColumns cols;
  int64_t n;
  B2_RETURN_NOT_OK(grouper_columns(g, keys, &cols, &n));
  if (n > 0xfffffff0ll) return set_error(B2_NOT_IMPLEMENTED, "grouper: batches above 2^32 rows must be split");
  Temp ids(ctx, s);
  B2_RETURN_NOT_OK(ids.alloc(sizeof(uint32_t) * (size_t)n));
  if (n == 0) {
    fill_out(out_ids, B2_UINT32, 0, 0, nullptr, ids.release());
    return B2_OK;
  }
  const int grid = grid_for(n, kBlock * 4, kSMs * 8);

  // direct-addressed mode: fit the window to this batch, or leave the mode for good
  if (insert && g->direct_eligible) {
    ScalarSlot mm(ctx);
    B2_RETURN_NOT_OK(mm.zero(s));
    B2_CUDA(cudaMemsetAsync(mm.dev(), 0xff, 8, s));
    grouper_minmax_kernel<<<grid, kBlock, 0, s>>>(cols.data[0], g->layout.width[0], cols.valid[0], n, 
                                                  grouper_sign_flip(g->key_types[0]),
                                                  reinterpret_cast<unsigned long long*>(mm.dev()));
    B2_LAUNCHED();
    B2_RETURN_NOT_OK(mm.fetch(s));
    uint64_t mn = static_cast<uint64_t>(mm.host()[0]), mx = static_cast<uint64_t>(mm.host()[1]);
    if (mn <= mx) {
      if (mn < g->obs_min) g->obs_min = mn;
      if (mx > g->obs_max) g->obs_max = mx;
    }
    const bool seen_any = g->obs_min <= g->obs_max;
    const uint64_t lo_need = seen_any ? g->obs_min : 0, hi_need = seen_any ? g->obs_max : 0;
    const uint64_t width = hi_need - lo_need;  // span - 1
    const uint64_t roomy = (uint64_t)n * 8 > 65536 ? (uint64_t)n * 8 : 65536;
    bool fits = width < kDirectMaxRange && (g->direct || width < roomy);
    if (fits) {
      if (!g->direct || lo_need < g->dt.lo || hi_need - g->dt.lo >= g->dt.cap)
        B2_RETURN_NOT_OK(grouper_build_direct(g, lo_need, hi_need, s));
    } else {
      const bool was_direct = g->direct;
      grouper_free_direct(g, s);
      g->direct_eligible = false;
      if (was_direct) B2_RETURN_NOT_OK(grouper_rebuild(g, next_pow2((uint64_t)g->num_groups * 4 + 1024), s));
    }
  }

  if (g->direct) {
    const void* kd = cols.data[0];
    const int kw = g->layout.width[0];
    if (!insert) {
      Temp bits(ctx, s);
      B2_RETURN_NOT_OK(bits.alloc(bitmap_alloc_bytes(n)));
      B2_CUDA(cudaMemsetAsync(bits.ptr, 0, bitmap_alloc_bytes(n), s));
      ScalarSlot slot(ctx);
      B2_RETURN_NOT_OK(slot.zero(s));
      grouper_direct_lookup_kernel<<<grid, kBlock, 0, s>>>(kd, kw, cols.valid[0], n, g->dt, ids.as<uint32_t>(),
                                                           bits.as<uint32_t>(), slot.dev());
      B2_LAUNCHED();
      B2_RETURN_NOT_OK(slot.fetch(s));
      int64_t nulls = n - slot.host()[0];
      fill_out(out_ids, B2_UINT32, n, nulls, nulls ? bits.release() : nullptr, ids.release());
      return B2_OK;
    }
    const bool fresh = g->num_groups == 0;
    if (fresh) grouper_direct_first_kernel<true><<<grid, kBlock, 0, s>>>(kd, kw, cols.valid[0], n, g->dt, ids.as<uint32_t>());
    else grouper_direct_first_kernel<false><<<grid, kBlock, 0, s>>>(kd, kw, cols.valid[0], n, g->dt, ids.as<uint32_t>());
    B2_LAUNCHED();
    Temp flags(ctx, s);
    B2_RETURN_NOT_OK(flags.alloc(bitmap_alloc_bytes(n)));
    B2_CUDA(cudaMemsetAsync(flags.ptr, 0, bitmap_alloc_bytes(n), s));
    grouper_direct_flag_kernel<<<grid_for((int64_t)g->dt.cap + 1, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(
        g->dt, flags.as<uint32_t>());
    B2_LAUNCHED();
    FilterBitmaps fb;
    fb.mask_data = BitmapReader(flags.ptr, 0, n);
    fb.mask_valid = BitmapReader(nullptr, 0, n);
    fb.values_valid = BitmapReader(nullptr, 0, n);
    fb.emit_null = 0;
    Temp offsets(ctx, s);
    int64_t n_new = 0, unused = 0;
    B2_RETURN_NOT_OK(filter_plan(ctx, fb, n, false, &offsets, &n_new, &unused, s));
    if ((uint64_t)g->num_groups + (uint64_t)n_new >= kNoId)
      return set_error(B2_CAPACITY_ERROR, "grouper: more than 2^32-1 groups");
    if (n_new > 0) {
      B2_RETURN_NOT_OK(grouper_reserve_uniques(g, (uint64_t)g->num_groups + n_new, s));
      grouper_assign_kernel<true><<<(unsigned)tiles_for(n), kBlock, 0, s>>>(
          g->layout, cols, n, g->table, g->dt, nullptr, fb.mask_data, offsets.as<int64_t>(), g->num_groups,
          g->uniq_keys, g->uniq_null);
      B2_LAUNCHED();
      g->num_groups += static_cast<uint32_t>(n_new);
      grouper_direct_gather_kernel<<<grid, kBlock, 0, s>>>(kd, kw, cols.valid[0], n, g->dt, ids.as<uint32_t>());
      B2_LAUNCHED();
      if (out_new_rows) B2_RETURN_NOT_OK(grouper_new_rows(ctx, flags.ptr, n, out_new_rows, s));
    }
    fill_out(out_ids, B2_UINT32, n, 0, nullptr, ids.release());
    return B2_OK;
  }

  Temp row_slot(ctx, s);
  B2_RETURN_NOT_OK(row_slot.alloc(sizeof(uint32_t) * (size_t)n));

  if (!insert) {
    if (g->cap == 0) B2_RETURN_NOT_OK(grouper_rebuild(g, 1024, s));
    Temp bits(ctx, s);
    B2_RETURN_NOT_OK(bits.alloc(bitmap_alloc_bytes(n)));
    B2_CUDA(cudaMemsetAsync(bits.ptr, 0, bitmap_alloc_bytes(n), s));
    ScalarSlot slot(ctx);
    B2_RETURN_NOT_OK(slot.zero(s));
    grouper_lookup_kernel<<<grid, kBlock, 0, s>>>(g->layout, cols, n, g->table, ids.as<uint32_t>(), bits.as<uint32_t>(),
                                                  slot.dev());
    B2_LAUNCHED();
    B2_RETURN_NOT_OK(slot.fetch(s));
    int64_t nulls = n - slot.host()[0];
    fill_out(out_ids, B2_UINT32, n, nulls, nulls ? bits.release() : nullptr, ids.release());
    return B2_OK;
  }

  // make room: at most n new groups; keep load <= 1/2 with a modest first guess
  if (g->cap == 0) {
    uint64_t guess = next_pow2((uint64_t)(n < 512 ? 1024 : (n < (1 << 22) ? 2 * n : (1 << 23))));
    B2_RETURN_NOT_OK(grouper_rebuild(g, guess, s));
  }
  while (true) {
    ScalarSlot slot(ctx);
    B2_RETURN_NOT_OK(slot.zero(s));
    grouper_insert_kernel<<<grid, kBlock, 0, s>>>(g->layout, cols, n, g->table, row_slot.as<uint32_t>(),
                                                  ids.as<uint32_t>(), slot.dev());
    B2_LAUNCHED();
    B2_RETURN_NOT_OK(slot.fetch(s));
    if (!slot.host()[0]) break;
    // probe limit hit: grow 4x; keys without ids are dropped by the rebuild and re-inserted
    B2_RETURN_NOT_OK(grouper_rebuild(g, g->cap * 4, s));
  }
  Temp flags(ctx, s);
  B2_RETURN_NOT_OK(flags.alloc(bitmap_alloc_bytes(n)));
  B2_CUDA(cudaMemsetAsync(flags.ptr, 0, bitmap_alloc_bytes(n), s));
  if (g->cap <= (uint64_t)n) {
    grouper_flag_slots_kernel<<<grid_for((int64_t)g->cap + 2, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(g->table,
                                                                                                    flags.as<uint32_t>());
  } else {
    grouper_flag_rows_kernel<<<grid, kBlock, 0, s>>>(n, g->table, row_slot.as<uint32_t>(), flags.as<uint32_t>());
  }
  B2_LAUNCHED();
  FilterBitmaps fb;
  fb.mask_data = BitmapReader(flags.ptr, 0, n);
  fb.mask_valid = BitmapReader(nullptr, 0, n);
  fb.values_valid = BitmapReader(nullptr, 0, n);
  fb.emit_null = 0;
  Temp offsets(ctx, s);
  int64_t n_new = 0, unused = 0;
  B2_RETURN_NOT_OK(filter_plan(ctx, fb, n, false, &offsets, &n_new, &unused, s));
  if ((uint64_t)g->num_groups + (uint64_t)n_new >= kNoId)
    return set_error(B2_CAPACITY_ERROR, "grouper: more than 2^32-1 groups");
  if (n_new > 0) {
    B2_RETURN_NOT_OK(grouper_reserve_uniques(g, (uint64_t)g->num_groups + n_new, s));
    grouper_assign_kernel<false><<<(unsigned)tiles_for(n), kBlock, 0, s>>>(g->layout, cols, n, g->table, g->dt,
                                                                   row_slot.as<uint32_t>(), fb.mask_data,
                                                                   offsets.as<int64_t>(), g->num_groups,
                                                                   g->uniq_keys, g->uniq_null);
    B2_LAUNCHED();
    g->num_groups += static_cast<uint32_t>(n_new);
  }
  if (n_new > 0) {  // otherwise the probe resolved every row
    grouper_gather_kernel<<<grid, kBlock, 0, s>>>(n, g->table, row_slot.as<uint32_t>(), ids.as<uint32_t>());
    B2_LAUNCHED();
    if (out_new_rows) B2_RETURN_NOT_OK(grouper_new_rows(ctx, flags.ptr, n, out_new_rows, s));
  }
  fill_out(out_ids, B2_UINT32, n, 0, nullptr, ids.release());
  // keep load factor <= 1/2 for the next batch
  if ((uint64_t)g->num_groups * 2 > g->cap) {
    B2_CUDA(cudaStreamSynchronize(s));
    B2_RETURN_NOT_OK(grouper_rebuild(g, next_pow2((uint64_t)g->num_groups * 4), s));
  }
  return B2_OK;
}

namespace b2 {
// internal to the library (grouper_wide.cu): keys that are hashes never fit a direct-addressed window -- skip the probing
void grouper_never_direct(B2Grouper* g) {
  g->never_direct = true;
  g->direct_eligible = false;
}
// Consume that also returns the rows at which this batch's new groups first occur (in id order)
int grouper_consume_new_rows(B2Grouper* g, const B2Array* keys, B2Array* out_ids, B2Array* out_new_rows, cudaStream_t s) {
  B2_CUDA(cudaSetDevice(g->ctx->device));
  return grouper_run(g, keys, out_ids, true, s, out_new_rows);
}
}  // namespace b2

extern "C" {

int b2_grouper_create(B2Context* ctx, const int32_t* key_types, int n_keys, B2Grouper** out) {
  if (!ctx || !key_types || !out) return set_error(B2_INVALID, "b2_grouper_create: null argument");
  if (n_keys < 1 || n_keys > kMaxKeys) return set_error(B2_NOT_IMPLEMENTED, "grouper: 1..%d key columns supported", kMaxKeys);
  KeyLayout L{};
  L.n_keys = n_keys;
  int bits = 0;
  {
    bool wide = false;
    int total = 0;
    for (int j = 0; j < n_keys; ++j) {
      wide = wide || type_is_binary_like(key_types[j]);
      total += 8 * type_width(key_types[j]) + (n_keys > 1 ? 1 : 0);
    }
    if (wide || total > 64) {
      WideGrouper* w = nullptr;
      B2_RETURN_NOT_OK(wide_create(ctx, key_types, n_keys, &w));
      B2Grouper* g = new B2Grouper();
      g->ctx = ctx;
      g->wide = w;
      g->layout.n_keys = n_keys;
      *out = g;
      return B2_OK;
    }
  }
  for (int j = 0; j < n_keys; ++j) {
    int w = type_width(key_types[j]);
    if (w == 0) return set_error(B2_NOT_IMPLEMENTED, "grouper: key type id %d is not fixed-width", key_types[j]);
    L.width[j] = w;
    L.bit_off[j] = bits;
    bits += 8 * w;
    L.null_bit[j] = -1;
  }
  if (n_keys > 1) {
    for (int j = 0; j < n_keys; ++j) L.null_bit[j] = bits++;
  }
  if (bits > 64)
    return set_error(B2_NOT_IMPLEMENTED, "grouper: encoded key needs %d bits; at most 64 are supported", bits);
  B2Grouper* g = new B2Grouper();
  g->ctx = ctx;
  g->layout = L;
  for (int j = 0; j < n_keys; ++j) g->key_types[j] = key_types[j];
  g->direct_eligible = n_keys == 1 && grouper_type_is_integer(key_types[0]) && grouper_direct_enabled();
  *out = g;
  return B2_OK;
}

void b2_grouper_destroy(B2Grouper* g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaStream_t s = g->ctx->stream;
  if (g->wide) wide_destroy(g->wide);
  grouper_free_table(g, s);
  grouper_free_direct(g, s);
  if (g->uniq_keys) g->ctx->free(g->uniq_keys, s);
  if (g->uniq_null) g->ctx->free(g->uniq_null, s);
  delete g;
}

int b2_grouper_consume(B2Grouper* g, const B2Array* keys, B2Array* out_ids, void* stream) {
  if (!g || !out_ids) return set_error(B2_INVALID, "b2_grouper_consume: null argument");
  B2_CUDA(cudaSetDevice(g->ctx->device));
  if (g->wide) return wide_run(g->wide, keys, out_ids, true, g->ctx->pick(stream));
  return grouper_run(g, keys, out_ids, true, g->ctx->pick(stream));
}

int b2_grouper_lookup(B2Grouper* g, const B2Array* keys, B2Array* out_ids, void* stream) {
  if (!g || !out_ids) return set_error(B2_INVALID, "b2_grouper_lookup: null argument");
  B2_CUDA(cudaSetDevice(g->ctx->device));
  if (g->wide) return wide_run(g->wide, keys, out_ids, false, g->ctx->pick(stream));
  return grouper_run(g, keys, out_ids, false, g->ctx->pick(stream));
}

int b2_grouper_num_groups(const B2Grouper* g, uint32_t* out) {
  if (!g || !out) return set_error(B2_INVALID, "b2_grouper_num_groups: null argument");
  *out = g->wide ? wide_num_groups(g->wide) : g->num_groups;
  return B2_OK;
}

int b2_grouper_uniques(B2Grouper* g, B2Array* out_keys, void* stream) {
  if (!g || !out_keys) return set_error(B2_INVALID, "b2_grouper_uniques: null argument");
  B2Context* ctx = g->ctx;
  cudaStream_t s = ctx->pick(stream);
  B2_CUDA(cudaSetDevice(ctx->device));
  if (g->wide) return wide_uniques(g->wide, out_keys, s);
  const uint32_t n = g->num_groups;
  for (int j = 0; j < g->layout.n_keys; ++j) {
    Temp data(ctx, s), bits(ctx, s);
    B2_RETURN_NOT_OK(data.alloc((size_t)n * g->layout.width[j]));
    int64_t nulls = 0;
    if (n > 0) {
      B2_RETURN_NOT_OK(bits.alloc(bitmap_alloc_bytes(n)));
      B2_CUDA(cudaMemsetAsync(bits.ptr, 0, bitmap_alloc_bytes(n), s));
      ScalarSlot slot(ctx);
      B2_RETURN_NOT_OK(slot.zero(s));
      grouper_decode_kernel<<<grid_for(n, kBlock * 4, kSMs * 8), kBlock, 0, s>>>(
          g->layout, j, g->uniq_keys, g->uniq_null, n, data.ptr, bits.as<uint32_t>(), slot.dev());
      B2_LAUNCHED();
      B2_RETURN_NOT_OK(slot.fetch(s));
      nulls = (int64_t)n - slot.host()[0];
    }
    fill_out(&out_keys[j], g->key_types[j], n, nulls, nulls ? bits.release() : nullptr, data.release());
  }
  return B2_OK;
}

int b2_grouper_reset(B2Grouper* g) {
  if (!g) return set_error(B2_INVALID, "b2_grouper_reset: null argument");
  B2_CUDA(cudaSetDevice(g->ctx->device));
  if (g->wide) return wide_reset(g->wide);
  cudaStream_t s = g->ctx->stream;
  grouper_free_table(g, s);
  grouper_free_direct(g, s);
  g->direct_eligible = !g->never_direct && g->layout.n_keys == 1 && grouper_type_is_integer(g->key_types[0]) && grouper_direct_enabled();
  g->obs_min = ~0ull;
  g->obs_max = 0;
  g->num_groups = 0;
  return B2_OK;
}

}  // extern "C"
