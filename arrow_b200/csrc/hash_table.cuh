// hash_table.cuh -- the open-addressing key table shared by the Grouper and the fused
// group-by.  B200-native stand-in for the reference's SwissTable + row table + Hashing32
// (cpp/src/arrow/compute/key_map_internal.h:41-90, key_hash_internal.h:38-67,
// row/encode_internal.h): keys of up to 64 encoded bits are stored directly in the slot
// array, claimed with one 64-bit atomicCAS, probed linearly.
#pragma once
#include "common.cuh"

namespace b2 {

// groupby_fused.cu: sums[id] += value, counts[id] += 1 over the valid values through the dense packed-state path
int dense_sum_count_by_id(B2Context* ctx, const uint32_t* ids, uint64_t num_groups, const void* values, int value_type,
                          BitmapReader val_valid, int64_t n, unsigned long long* sums, unsigned long long* counts,
                          cudaStream_t s, bool* done);

constexpr uint64_t kEmptyKey = 0xffffffffffffffffull;
constexpr int kMaxProbe = 256;

// murmur3 fmix64: full-avalanche mix of the encoded key
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// Find or claim the slot of `key` in keys[0..cap) (cap = mask+1, power of two; `stride`
// = slot pitch in uint64 units).  The two keys that cannot live in the open-addressing
// area -- the bit pattern used as the empty marker and the NULL key -- own the two extra
// slots cap and cap+1.  Returns -1 when the probe limit is hit (table too full).
__device__ __forceinline__ int64_t table_find_or_insert(unsigned long long* keys, uint64_t mask, int stride,
                                                        uint64_t key, bool is_null, bool* inserted) {
  *inserted = false;
  if (is_null || key == kEmptyKey) {
    uint64_t slot = mask + 1 + (is_null ? 1 : 0);
    unsigned long long* p = keys + slot * stride;
    // claim marker for the special slots: 0 = taken (they start as kEmptyKey)
    if (*reinterpret_cast<volatile unsigned long long*>(p) != 0ull) {
      unsigned long long old = atomicCAS(p, (unsigned long long)kEmptyKey, 0ull);
      *inserted = (old == kEmptyKey);
    }
    return static_cast<int64_t>(slot);
  }
  uint64_t slot = hash64(key) & mask;
  for (int probe = 0; probe < kMaxProbe; ++probe) {
    unsigned long long* p = keys + slot * stride;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
    if (cur == key) return static_cast<int64_t>(slot);
    if (cur == kEmptyKey) {
      unsigned long long old = atomicCAS(p, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) {
        *inserted = true;
        return static_cast<int64_t>(slot);
      }
      if (old == key) return static_cast<int64_t>(slot);
    }
    slot = (slot + 1) & mask;
  }
  return -1;
}

__device__ __forceinline__ int64_t table_find(const unsigned long long* keys, uint64_t mask, int stride,
                                              uint64_t key, bool is_null) {
  if (is_null || key == kEmptyKey) {
    uint64_t slot = mask + 1 + (is_null ? 1 : 0);
    return keys[slot * stride] == 0ull ? static_cast<int64_t>(slot) : -1;
  }
  uint64_t slot = hash64(key) & mask;
  for (int probe = 0; probe < kMaxProbe; ++probe) {
    unsigned long long cur = keys[slot * stride];
    if (cur == key) return static_cast<int64_t>(slot);
    if (cur == kEmptyKey) return -1;
    slot = (slot + 1) & mask;
  }
  return -1;
}

// load one fixed-width key value zero-extended to 64 bits
__device__ __forceinline__ uint64_t load_key_bits(const void* data, int width, int64_t i) {
  switch (width) {
    case 1: return static_cast<const uint8_t*>(data)[i];
    case 2: return static_cast<const uint16_t*>(data)[i];
    case 4: return static_cast<const uint32_t*>(data)[i];
    default: return static_cast<const uint64_t*>(data)[i];
  }
}

inline uint64_t next_pow2(uint64_t v) {
  uint64_t r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace b2
