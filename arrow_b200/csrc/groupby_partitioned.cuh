// groupby_partitioned.cuh -- the high-cardinality path of the fused group-by.
//
// Global atomics cannot carry config 3 (1B rows, 10M groups): every row is a random
// 32-byte-sector read-modify-write in a table far larger than L2, and SIMT linear probing
// runs as long as the slowest lane of a warp (ncu: 197 thread-instructions and 129 DRAM
// bytes per row, 89 % long-scoreboard stalls).  This path makes the rows of one key meet
// in SHARED memory instead:
//   1. part_hist_kernel  : histogram of the low hash digits of every key (one read of the keys)
//   2. part_pass_kernel  : 1 or 2 LSD radix passes on hash64(key) digits (the SortIndices
//                          onesweep machinery: ticketed tiles, smem-atomic ranking, decoupled
//                          look-back, smem-staged contiguous writes) move (key, value, flags)
//                          tuples so that rows with equal hash digits become contiguous --
//                          256 or 65536 partitions, i.e. ~G/65536 distinct keys per partition;
//   3. preagg_kernel     : each CTA streams 8192-row slices of the partitioned tuples through
//                          a 2048-slot shared-memory table (smem CAS + smem atomicAdd), then
//                          flushes the few distinct (key, sum, count) entries it found into
//                          the global table -- G x few global atomics instead of 3 per row.
// The global table stays the single source of truth (robust to skew: a heavy key just
// flushes once per slice; a slice with too many keys spills rows straight to the table).
// Traffic (2 passes): 8 + 2*(17+17) + 17 = 93 B/row streamed, vs ~160 B/row random before.
#pragma once
#include <type_traits>

#include "bitmap.h"
#include "hash_table.cuh"

namespace b2 {

constexpr int kPartThreads = 512;
constexpr int kPartWarps = kPartThreads / 32;
constexpr int kPartItems = 8;
constexpr int kPartTile = kPartThreads * kPartItems;  // 4096 rows per tile (2048 measured 20 % slower: shorter write runs, 2x look-back)
constexpr int kPartRadix = 256;
constexpr uint32_t kPFlagAgg = 1u << 30, kPFlagIncl = 2u << 30, kPValMask = (1u << 30) - 1u;

struct Tuples {
  unsigned long long* keys;
  unsigned long long* vals;  // int64 / uint64 / double bits
  uint8_t* flags;            // bit0 = value valid, bit1 = key is null
};

template <typename V>
__device__ __forceinline__ unsigned long long value_bits(V v) {
  if constexpr (std::is_floating_point<V>::value) return static_cast<unsigned long long>(__double_as_longlong(static_cast<double>(v)));
  else if constexpr (std::is_signed<V>::value) return static_cast<unsigned long long>(static_cast<long long>(v));
  else return static_cast<unsigned long long>(v);
}

struct RawColumns {
  const void* keys;    // advanced by offset * KW
  const void* values;  // advanced by offset * sizeof(V)
  BitmapReader key_valid, val_valid;
  int64_t row0;        // first row of this chunk
};

__device__ __forceinline__ unsigned part_digit(unsigned long long key, unsigned flags, int shift) {
  return (flags & 2u) ? 0u : static_cast<unsigned>(hash64(key) >> shift) & (kPartRadix - 1);
}

template <int KW>
__global__ void __launch_bounds__(kBlock) part_hist_kernel(RawColumns c, int64_t n, int passes, unsigned long long* hist) {
  __shared__ uint32_t s_hist[2 * kPartRadix];
  for (int i = threadIdx.x; i < 2 * kPartRadix; i += kBlock) s_hist[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const int64_t r = c.row0 + i;
    const unsigned f = c.key_valid.bit(r) ? 0u : 2u;
    const unsigned long long k = f ? 0ull : load_key_bits(c.keys, KW, r);
    atomicAdd(&s_hist[part_digit(k, f, 0)], 1u);
    if (passes > 1) atomicAdd(&s_hist[kPartRadix + part_digit(k, f, 8)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kPartRadix; i += kBlock)
    if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

__global__ void __launch_bounds__(kPartRadix) part_scan_kernel(const unsigned long long* hist, uint32_t* digit_base) {
  __shared__ uint32_t s[kPartRadix];
  const int p = blockIdx.x, d = threadIdx.x;
  s[d] = static_cast<uint32_t>(hist[p * kPartRadix + d]);
  __syncthreads();
  if (d == 0) {
    uint32_t run = 0;
    for (int i = 0; i < kPartRadix; ++i) {
      uint32_t v = s[i];
      s[i] = run;
      run += v;
    }
  }
  __syncthreads();
  digit_base[p * kPartRadix + d] = s[d];
}

struct PartArgs {
  RawColumns raw;   // FIRST pass input
  Tuples in, out;   // later passes read `in`
  uint32_t n;
  int shift;
  const uint32_t* digit_base;  // [256] exclusive bin offsets of this pass
  uint32_t* lookback;          // [n_tiles][256], zeroed, followed by the ticket
  uint32_t* ticket;
};

constexpr size_t part_smem_bytes() {
  return kPartTile * 8 * 2 + kPartTile + 3 * kPartRadix * 4;
}

template <bool FIRST, typename V, int KW>
__global__ void __launch_bounds__(kPartThreads, 2) part_pass_kernel(PartArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem);
  unsigned long long* s_vals = s_keys + kPartTile;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_vals + kPartTile);  // [256] rows per bin in this tile
  uint32_t* s_bin = s_cnt + kPartRadix;
  uint32_t* s_gbase = s_bin + kPartRadix;
  uint8_t* s_flags = reinterpret_cast<uint8_t*>(s_gbase + kPartRadix);
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_warp_tot[kPartRadix / 32];

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
  for (int i = tid; i < kPartRadix; i += kPartThreads) s_cnt[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kPartTile;
  const uint32_t tile_n = (a.n - base) < (uint32_t)kPartTile ? (a.n - base) : (uint32_t)kPartTile;

  unsigned long long key[kPartItems], val[kPartItems];
  unsigned flg[kPartItems];
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const uint32_t i = warp * (32 * kPartItems) + j * 32 + lane;
    key[j] = ~0ull;
    val[j] = 0;
    flg[j] = 4u;  // padding marker
    if (i < tile_n) {
      if (FIRST) {
        const int64_t r = a.raw.row0 + base + i;
        unsigned f = a.raw.key_valid.bit(r) ? 0u : 2u;
        key[j] = f ? 0ull : load_key_bits(a.raw.keys, KW, r);
        if (a.raw.val_valid.bit(r)) {
          f |= 1u;
          val[j] = value_bits<V>(static_cast<const V*>(a.raw.values)[r]);
        }
        flg[j] = f;
      } else {
        key[j] = __ldcs(a.in.keys + base + i);
        val[j] = __ldcs(a.in.vals + base + i);
        flg[j] = a.in.flags[base + i];
      }
    }
  }
  // Rank inside the tile's bin with ONE returning shared-memory atomic per row
  // (ATOMS.ADD.u32: 0.17 cycles/lane/SM measured, vs 1.83 for MATCH.ANY -- profiles/smem_probe_r01.txt).
  // The order of rows inside a partition is irrelevant for grouping, so no stable ranking is needed.
  uint32_t rank_dig[kPartItems];  // rank inside the bin | digit << 16 (one register per row)
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const unsigned d = part_digit(key[j], flg[j], a.shift);
    const uint32_t r = (flg[j] & 4u) ? 0u : atomicAdd(&s_cnt[d], 1u);  // padding rows are neither counted nor staged
    rank_dig[j] = r | (d << 16);
  }
  __syncthreads();

  uint32_t run = 0, incl = 0;
  if (tid < kPartRadix) {
    run = s_cnt[tid];
    volatile uint32_t* lb = a.lookback;
    if (tile == 0) lb[tid] = kPFlagIncl | run;
    else lb[(size_t)tile * kPartRadix + tid] = kPFlagAgg | run;
    incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
  }
  __syncthreads();
  uint32_t bin_off = 0;
  if (tid < kPartRadix) {
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < kPartRadix / 32; ++w)
      if (w < (int)warp) woff += s_warp_tot[w];
    bin_off = woff + incl - run;
    s_bin[tid] = bin_off;
  }
  __syncthreads();
  // stage the tile in bin order first: it needs only tile-local offsets, and it gives the
  // predecessors time to publish before the look-back below has to wait for them
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    if (flg[j] & 4u) continue;
    const uint32_t pos = s_bin[rank_dig[j] >> 16] + (rank_dig[j] & 0xffffu);
    s_keys[pos] = key[j];
    s_vals[pos] = val[j];
    s_flags[pos] = static_cast<uint8_t>(flg[j]);
  }
  if (tid < kPartRadix) {
    uint32_t excl = 0;
    if (tile > 0) {
      excl = lookback_exclusive(a.lookback + tid, tile, kPartRadix);
      reinterpret_cast<volatile uint32_t*>(a.lookback)[(size_t)tile * kPartRadix + tid] = kPFlagIncl | (excl + run);
    }
    s_gbase[tid] = a.digit_base[tid] + excl - bin_off;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPartItems; ++j) {
    const uint32_t p = j * kPartThreads + tid;
    if (p < tile_n) {
      const unsigned long long k = s_keys[p];
      const unsigned f = s_flags[p];
      const uint32_t dst = s_gbase[part_digit(k, f, a.shift)] + p;
      a.out.keys[dst] = k;
      a.out.vals[dst] = s_vals[p];
      a.out.flags[dst] = static_cast<uint8_t>(f);
    }
  }
}

// ---- pre-aggregation in shared memory, flush to the global table ----
constexpr int kPreSlots = 2048;
constexpr int kPreSlice = 8192;  // rows per CTA iteration
constexpr int kPreProbe = 48;

struct FusedTableRef {
  unsigned long long* slots;
  uint64_t mask;
  // entries that met a full neighbourhood (probe limit) are parked here and replayed by the host
  // after it has grown the table: {key, sum bits} pairs + {count | key_null << 31}
  unsigned long long* ovf_pairs;
  unsigned int* ovf_counts;
  unsigned long long ovf_cap;  // entries the parking area can hold; counters[2] counts what did not fit
};

template <bool IS_FLOAT>
__device__ __forceinline__ void global_accumulate(const FusedTableRef& t, unsigned long long key, bool key_null,
                                                  unsigned long long sum_bits, unsigned count, unsigned long long* counters) {
  bool inserted;
  int64_t slot = table_find_or_insert(t.slots, t.mask, 4, key, key_null, &inserted);
  if (slot < 0) {
    const unsigned long long i = atomicAdd(&counters[0], 1ull);
    if (i < t.ovf_cap) {
      t.ovf_pairs[2 * i] = key;
      t.ovf_pairs[2 * i + 1] = sum_bits;
      t.ovf_counts[i] = count | (key_null ? 0x80000000u : 0u);
    } else {
      atomicAdd(&counters[2], 1ull);  // reported by the host as a capacity error
    }
    return;
  }
  if (inserted) atomicAdd(&counters[1], 1ull);
  if (count) {
    unsigned long long* p = t.slots + slot * 4;
    if (IS_FLOAT) atomicAdd(reinterpret_cast<double*>(p + 1), __longlong_as_double((long long)sum_bits));
    else atomicAdd(p + 1, sum_bits);
    atomicAdd(p + 2, (unsigned long long)count);
  }
}

template <bool RAW, bool IS_FLOAT, typename V, int KW>
__global__ void __launch_bounds__(kBlock) preagg_kernel(RawColumns raw, Tuples in, int64_t n, FusedTableRef table,
                                                        unsigned long long* counters) {
  __shared__ unsigned long long s_keys[kPreSlots + 2];  // +2: the empty-pattern key and the null key
  // integer sums live as (lo, hi) 32-bit halves updated with two native ATOMS.ADD.u32 and an explicit
  // carry: a 64-bit shared atomicAdd is a CAS loop on this part (0.64 cycles/lane spread, 40 when
  // contended on one key -- profiles/smem_probe_r01.txt); doubles keep the CAS-based atomicAdd
  __shared__ unsigned long long s_sums[kPreSlots + 2];
  __shared__ unsigned int s_counts[kPreSlots + 2];
  __shared__ uint8_t s_used[kPreSlots + 2];  // slot touched (a group can exist with count 0)
  const int64_t n_slices = (n + kPreSlice - 1) / kPreSlice;
  for (int64_t slice = blockIdx.x; slice < n_slices; slice += gridDim.x) {
    for (int i = threadIdx.x; i < kPreSlots + 2; i += kBlock) {
      s_keys[i] = kEmptyKey;
      s_sums[i] = 0;
      s_counts[i] = 0;
      s_used[i] = 0;
    }
    __syncthreads();
    const int64_t lo = slice * kPreSlice;
    const int64_t hi = lo + kPreSlice < n ? lo + kPreSlice : n;
    constexpr int kBatch = 8;  // rows loaded per thread before the dependent shared-memory work
    for (int64_t b0 = lo; b0 < hi; b0 += (int64_t)kBatch * kBlock) {
      unsigned long long kk[kBatch], vv[kBatch];
      unsigned ff[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int64_t i = b0 + u * kBlock + threadIdx.x;
        kk[u] = 0;
        vv[u] = 0;
        ff[u] = 8u;  // 8 = no row
        if (i < hi) {
          if (RAW) {
            const int64_t r = raw.row0 + i;
            unsigned f = raw.key_valid.bit(r) ? 0u : 2u;
            kk[u] = f ? 0ull : load_key_bits(raw.keys, KW, r);
            if (raw.val_valid.bit(r)) {
              f |= 1u;
              vv[u] = value_bits<V>(static_cast<const V*>(raw.values)[r]);
            }
            ff[u] = f;
          } else {
            kk[u] = __ldcs(in.keys + i);
            vv[u] = __ldcs(in.vals + i);
            ff[u] = in.flags[i];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned f = ff[u];
        if (f & 8u) continue;
        const unsigned long long k = kk[u], vb = vv[u];
        int slot = -1;
        if (f & 2u) {
          slot = kPreSlots + 1;
        } else if (k == kEmptyKey) {
          slot = kPreSlots;
        } else {
          unsigned s = static_cast<unsigned>(hash64(k) >> 20) & (kPreSlots - 1);
          for (int probe = 0; probe < kPreProbe; ++probe) {
            unsigned long long cur = s_keys[s];
            if (cur == k) {
              slot = s;
              break;
            }
            if (cur == kEmptyKey) {
              unsigned long long old = atomicCAS(&s_keys[s], (unsigned long long)kEmptyKey, k);
              if (old == kEmptyKey || old == k) {
                slot = s;
                break;
              }
            }
            s = (s + 1) & (kPreSlots - 1);
          }
        }
        if (slot < 0) {  // slice holds too many distinct keys: spill this row to the global table
          global_accumulate<IS_FLOAT>(table, k, false, vb, (f & 1u) ? 1u : 0u, counters);
          continue;
        }
        s_used[slot] = 1;
        if (f & 1u) {
          if (IS_FLOAT) {
            atomicAdd(reinterpret_cast<double*>(&s_sums[slot]), __longlong_as_double((long long)vb));
          } else {
            unsigned int* half = reinterpret_cast<unsigned int*>(&s_sums[slot]);  // little endian: [0] = lo, [1] = hi
            const unsigned int lo32 = static_cast<unsigned int>(vb), hi32 = static_cast<unsigned int>(vb >> 32);
            const unsigned int old = atomicAdd(half, lo32);
            const unsigned int carry = (old + lo32) < old ? 1u : 0u;
            if (hi32 + carry) atomicAdd(half + 1, hi32 + carry);
          }
          atomicAdd(&s_counts[slot], 1u);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPreSlots + 2; i += kBlock) {
      if (s_used[i]) {
        const bool knull = i == kPreSlots + 1;
        const unsigned long long k = i == kPreSlots ? kEmptyKey : s_keys[i];
        global_accumulate<IS_FLOAT>(table, knull ? 0ull : k, knull, s_sums[i], s_counts[i], counters);
      }
    }
    __syncthreads();
  }
}

template <bool IS_FLOAT>
__global__ void __launch_bounds__(kBlock) replay_overflow_kernel(FusedTableRef table, const unsigned long long* pairs,
                                                                 const unsigned int* counts, int64_t n,
                                                                 unsigned long long* counters) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    global_accumulate<IS_FLOAT>(table, pairs[2 * i], (counts[i] >> 31) != 0, pairs[2 * i + 1], counts[i] & 0x7fffffffu, counters);
}

}  // namespace b2
