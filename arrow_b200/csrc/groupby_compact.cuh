// groupby_compact.cuh -- the bandwidth-lean path of the fused group-by (hash_sum + hash_count).
//
// The partitioned path of groupby_partitioned.cuh moves 17-byte (key, value, flags) tuples twice and
// runs its passes at ~50 % of the HBM rate (profiles/partpass_prof_r01b: spills under the 64-register
// cap, barrier stalls, 2 CTAs/SM with nothing in flight while a tile is ranked).  This path attacks
// both the bytes and the stalls:
//
//   * 8-byte tuples.  The stats pass measures the exact key range [kmin, kmax] (order-preserving
//     encoding) and a sample of the value range; a row then travels as ONE 64-bit word
//         (key - kmin) << (vb + 1) | (value - vbase) << 1 | value_valid
//     whenever kb + vb + 1 <= 64 (kb, vb = bits of the two ranges).  The value window comes from the
//     type (<= 32-bit values) or from a sample (64-bit values) and is VERIFIED on every row in pass 1;
//     a row outside the window raises a flag and the host redoes the chunk on the general path, so the
//     result never depends on the sample.  Null-key rows never enter the tuples: pass 1 reduces them
//     into the null group's accumulator directly.
//   * Bulk-asynchronous loads (tma.cuh: cp.async.bulk + mbarrier, SASS UBLKCP).  Every pass is a
//     persistent kernel; while a CTA ranks / stages / writes tile i from registers, the copy engine is
//     already filling the input buffer with tile i+1, so there are always (CTAs per SM) x 32-64 KB of
//     reads in flight per SM without spending registers on them.
//   * Cheap partition hash (3 x 32-bit multiply + xorshift of the folded key) instead of murmur's
//     64-bit multiplies, recomputed from the tuple in every pass; the shared-memory pre-aggregation
//     table is indexed with the low bits of the same hash and, for narrow ranges, keeps 32-bit keys,
//     32-bit partial sums and 32-bit counts (12 bytes per slot, native ATOMS only).
//
// Traffic per row (2 passes): stats 8 + pass 1 (16.25 + 8) + pass 2 (8 + 8) + pre-aggregation 8
// = 56 B/row against 93 B/row before.  The global 32-byte-slot table of groupby_fused.cu stays the single
// source of truth, exactly as in the general path.
#pragma once
#include "groupby_partitioned.cuh"
#include "tma.cuh"

namespace b2 {

constexpr int kCThreads = 512;
constexpr int kCItems = 8;
constexpr int kCTile = kCThreads * kCItems;  // 4096 rows per tile
constexpr int kCWarps = kCThreads / 32;

struct CompactEnc {
  unsigned long long kmin;   // minimum of enc(key) over the non-null keys of the chunk
  unsigned long long kflip;  // enc(key) = zero-extended key bits ^ kflip (sign bit for signed key types)
  unsigned long long vbase;  // v' = sign/zero-extended value bits - vbase (mod 2^64), 0 <= v' < 2^vb
  int kb, vb;                // kb + vb + 1 <= 64
};

// partition hash of an encoded key: cheap, but every output bit depends on every input bit of both halves
__device__ __forceinline__ uint32_t compact_hash(unsigned long long enc_key) {
  uint32_t x = static_cast<uint32_t>(enc_key) ^ (static_cast<uint32_t>(enc_key >> 32) * 0x85EBCA77u);
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0xC2B2AE3Du;
  x ^= x >> 13;
  return x;
}

__device__ __forceinline__ unsigned long long compact_key(unsigned long long t, const CompactEnc& e) { return t >> (e.vb + 1); }

// ---- stats + histograms: one read of the key column ------------------------------------------------
struct CompactStats {
  unsigned long long kmin, kmax;      // of enc(key) over valid keys (kmin = ~0, kmax = 0 when there are none)
  unsigned long long null_keys;
  unsigned long long vmin, vmax;      // sampled value range in the value type's own order (int64 bits / uint64)
  unsigned long long sampled;
};

template <int KW>
__global__ void __launch_bounds__(kBlock) compact_stats_kernel(const void* __restrict__ keys, BitmapReader key_valid, int64_t row0,
                                                               int64_t n, unsigned long long kflip, int passes,
                                                               unsigned long long* __restrict__ hist, CompactStats* stats) {
  __shared__ uint32_t s_hist[2 * kPartRadix];
  __shared__ unsigned long long s_min[kWarpsPerBlock], s_max[kWarpsPerBlock], s_null[kWarpsPerBlock];
  for (int i = threadIdx.x; i < 2 * kPartRadix; i += kBlock) s_hist[i] = 0;
  __syncthreads();
  unsigned long long lo = ~0ull, hi = 0ull, nulls = 0;
  constexpr int U = 4;  // independent loads in flight per thread
  const int64_t stride = (int64_t)gridDim.x * kBlock * U;
  for (int64_t base = (int64_t)blockIdx.x * kBlock * U; base < n; base += stride) {
    unsigned long long k[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + u * kBlock + threadIdx.x;
      ok[u] = i < n;
      k[u] = ok[u] ? load_key_bits(keys, KW, row0 + i) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const int64_t i = base + u * kBlock + threadIdx.x;
      if (!key_valid.bit(row0 + i)) {
        ++nulls;
        continue;
      }
      const unsigned long long e = k[u] ^ kflip;
      lo = e < lo ? e : lo;
      hi = e > hi ? e : hi;
      const uint32_t h = compact_hash(e);
      atomicAdd(&s_hist[h >> 24], 1u);
      if (passes > 1) atomicAdd(&s_hist[kPartRadix + ((h >> 16) & 255u)], 1u);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_down_sync(0xffffffffu, lo, o), b = __shfl_down_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
    nulls += __shfl_down_sync(0xffffffffu, nulls, o);
  }
  if (lane_id() == 0) {
    s_min[threadIdx.x >> 5] = lo;
    s_max[threadIdx.x >> 5] = hi;
    s_null[threadIdx.x >> 5] = nulls;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kWarpsPerBlock; ++w) {
      lo = s_min[w] < lo ? s_min[w] : lo;
      hi = s_max[w] > hi ? s_max[w] : hi;
      nulls += s_null[w];
    }
    atomicMin(&stats->kmin, lo);
    atomicMax(&stats->kmax, hi);
    if (nulls) atomicAdd(&stats->null_keys, nulls);
  }
  for (int i = threadIdx.x; i < passes * kPartRadix; i += kBlock)
    if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// sign- or zero-extended 64-bit image of a value of width vw
__device__ __forceinline__ unsigned long long load_value_bits(const void* p, int vw, bool vsigned, int64_t i) {
  switch (vw) {
    case 1: return vsigned ? (unsigned long long)(long long)static_cast<const int8_t*>(p)[i] : static_cast<const uint8_t*>(p)[i];
    case 2: return vsigned ? (unsigned long long)(long long)static_cast<const int16_t*>(p)[i] : static_cast<const uint16_t*>(p)[i];
    case 4: return vsigned ? (unsigned long long)(long long)static_cast<const int32_t*>(p)[i] : static_cast<const uint32_t*>(p)[i];
    default: return static_cast<const unsigned long long*>(p)[i];
  }
}

// min / max of <= 65536 evenly spaced valid 64-bit values (the window is verified on every row later)
__global__ void __launch_bounds__(kBlock) compact_value_sample_kernel(const unsigned long long* __restrict__ vals, BitmapReader val_valid,
                                                                      int64_t row0, int64_t n, int64_t step, bool vsigned,
                                                                      CompactStats* stats) {
  const unsigned long long flip = vsigned ? 0x8000000000000000ull : 0ull;  // order-preserving for the comparison only
  for (int64_t s = blockIdx.x * (int64_t)kBlock + threadIdx.x; s * step < n; s += (int64_t)gridDim.x * kBlock) {
    const int64_t i = row0 + s * step;
    if (!val_valid.bit(i)) continue;
    const unsigned long long v = vals[i] ^ flip;
    atomicMin(&stats->vmin, v);
    atomicMax(&stats->vmax, v);
    atomicAdd(&stats->sampled, 1ull);
  }
}

// ---- the partition passes ---------------------------------------------------------------------------
struct CompactArgs {
  // pass 1 input: the user's columns, advanced to the first row of the chunk
  const uint8_t* keys;
  const uint8_t* vals;
  int kw, vw;
  bool vsigned, bulk_ok;
  BitmapReader key_valid, val_valid;  // addressed with row0 + i
  int64_t row0;
  // later passes read tuples
  const unsigned long long* in;
  unsigned long long* out;
  uint32_t n;  // rows of the chunk (pass 1) / tuples (pass 2)
  CompactEnc enc;
  int shift;                     // 24: first hash digit, 16: second
  const uint32_t* digit_base;    // [256] exclusive bin offsets of this pass
  uint32_t* lookback;            // [n_tiles][256], zeroed, followed by the ticket
  uint32_t* ticket;
  unsigned long long* null_acc;  // [0] sum bits, [1] count of valid values, [2] rows  (pass 1: rows with a null key)
  unsigned int* overflow;        // pass 1: set when a valid value lies outside the window
};

constexpr size_t compact_pass_smem(bool first) {
  return (first ? (size_t)kCTile * 16 : (size_t)kCTile * 8) + (size_t)kCTile * 8;
}

// KW / VW: key / value byte widths known at compile time for the common shapes (8,8) and (4,4); 0 = read a.kw / a.vw
template <bool FIRST, int KW, int VW>
__global__ void __launch_bounds__(kCThreads, 2) compact_pass_kernel(CompactArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* in_keys = smem;                                                       // FIRST: kCTile * kw bytes (<= 32 KB)
  uint8_t* in_vals = smem + (size_t)kCTile * 8;                                  // FIRST: kCTile * vw bytes
  unsigned long long* in_tuples = reinterpret_cast<unsigned long long*>(smem);   // !FIRST
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(smem + (FIRST ? (size_t)kCTile * 16 : (size_t)kCTile * 8));
  __shared__ uint32_t s_cnt[kPartRadix], s_bin[kPartRadix], s_gbase[kPartRadix];
  __shared__ uint32_t s_warp_tot[kPartRadix / 32];
  __shared__ uint32_t s_cur, s_total;
  __shared__ uint8_t s_dig[kCTile];  // digit of every staged tuple: the write-out does not recompute the hash
  __shared__ __align__(8) uint64_t s_bar;

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kw = KW ? KW : a.kw, vw = VW ? VW : a.vw;
  const uint32_t n_tiles = (a.n + kCTile - 1) / kCTile;
  const CompactEnc enc = a.enc;
  const unsigned long long vmask = (enc.vb >= 64) ? ~0ull : ((1ull << enc.vb) - 1ull);

  // issue the loads of one tile into the input buffer (thread 0: bulk copies; or everybody: plain loads)
  auto load_tile = [&](uint32_t tile) {
    const uint32_t base = tile * kCTile;
    const uint32_t tile_n = (a.n - base) < (uint32_t)kCTile ? (a.n - base) : (uint32_t)kCTile;
    if (FIRST) {
      const bool bulk = a.bulk_ok && tile_n == (uint32_t)kCTile;
      if (bulk) {
        if (tid == 0) {
          const uint32_t bk = (uint32_t)kCTile * kw, bv = (uint32_t)kCTile * vw;
          mbar_arrive_expect_tx(&s_bar, bk + bv);
          bulk_copy_g2s_chunked(in_keys, a.keys + (size_t)base * kw, bk, &s_bar);
          bulk_copy_g2s_chunked(in_vals, a.vals + (size_t)base * vw, bv, &s_bar);
        }
      } else {  // unaligned slice or the partial last tile: exact-bounds element copies by every thread
        for (uint32_t i = tid; i < tile_n; i += kCThreads) {
          switch (kw) {
            case 1: in_keys[i] = a.keys[(size_t)base + i]; break;
            case 2: reinterpret_cast<uint16_t*>(in_keys)[i] = reinterpret_cast<const uint16_t*>(a.keys)[(size_t)base + i]; break;
            case 4: reinterpret_cast<uint32_t*>(in_keys)[i] = reinterpret_cast<const uint32_t*>(a.keys)[(size_t)base + i]; break;
            default: reinterpret_cast<unsigned long long*>(in_keys)[i] = reinterpret_cast<const unsigned long long*>(a.keys)[(size_t)base + i]; break;
          }
          switch (vw) {
            case 1: in_vals[i] = a.vals[(size_t)base + i]; break;
            case 2: reinterpret_cast<uint16_t*>(in_vals)[i] = reinterpret_cast<const uint16_t*>(a.vals)[(size_t)base + i]; break;
            case 4: reinterpret_cast<uint32_t*>(in_vals)[i] = reinterpret_cast<const uint32_t*>(a.vals)[(size_t)base + i]; break;
            default: reinterpret_cast<unsigned long long*>(in_vals)[i] = reinterpret_cast<const unsigned long long*>(a.vals)[(size_t)base + i]; break;
          }
        }
        if (tid == 0) mbar_arrive_expect_tx(&s_bar, 0);  // completes the phase at once: the wait below stays uniform
      }
    } else {
      if (tid == 0) {
        // tuple buffers are pool allocations padded to whole tiles: the partial last tile may be read in full
        const uint32_t bytes = ((tile_n * 8u) + 15u) & ~15u;
        mbar_arrive_expect_tx(&s_bar, bytes);
        bulk_copy_g2s_chunked(in_tuples, a.in + base, bytes, &s_bar);
      }
    }
  };

  // tile tickets: the atomic's round trip (~1 us) must never sit between two barriers, so thread 0 always holds
  // the ticket AFTER the next one in a register (fetched a whole iteration before it is published)
  uint32_t ticket_ahead = 0;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    s_cur = atomicAdd(a.ticket, 1u);
    ticket_ahead = atomicAdd(a.ticket, 1u);
  }
  for (int i = tid; i < kPartRadix; i += kCThreads) s_cnt[i] = 0;
  __syncthreads();
  uint32_t tile = s_cur;
  if (tile < n_tiles) load_tile(tile);
  uint32_t parity = 0;
  unsigned long long null_sum = 0, null_cnt = 0, null_rows = 0;

  while (tile < n_tiles) {
    const uint32_t base = tile * kCTile;
    const uint32_t tile_n = (a.n - base) < (uint32_t)kCTile ? (a.n - base) : (uint32_t)kCTile;
    // validity bits of this thread's rows (row = j * 512 + tid): fetched before the wait so the loads overlap it
    uint32_t kvalid = 0xffu, vvalid = 0xffu;
    if (FIRST) {
      // lane j (< 8) loads the key-validity word of row group j, lane 8 + j the value-validity word; a shuffle hands
      // every lane its own bit (32 lanes calling word32 each cost 50 instructions per row)
      uint32_t word = 0xffffffffu;
      if (lane < 2 * kCItems) {
        const int j = lane & (kCItems - 1);
        const int64_t w = (a.row0 + base + j * kCThreads + warp * 32) >> 5;  // row0 and base are multiples of 32
        word = lane < kCItems ? a.key_valid.word32(w) : a.val_valid.word32(w);
      }
      kvalid = vvalid = 0;
#pragma unroll
      for (int j = 0; j < kCItems; ++j) {
        kvalid |= ((__shfl_sync(0xffffffffu, word, j) >> lane) & 1u) << j;
        vvalid |= ((__shfl_sync(0xffffffffu, word, kCItems + j) >> lane) & 1u) << j;
      }
    }
    mbar_wait(&s_bar, parity);
    parity ^= 1u;
    // bulk copies are visible to every thread that observed the barrier phase; only the plain-load fallback (unaligned
    // slice / partial last tile) needs a CTA barrier for the other threads' writes -- the condition is uniform per tile
    if (FIRST && !(a.bulk_ok && tile_n == (uint32_t)kCTile)) __syncthreads();

    unsigned long long t[kCItems];
    uint32_t rank_dig[kCItems];
#pragma unroll
    for (int j = 0; j < kCItems; ++j) {
      const uint32_t r = j * kCThreads + tid;
      rank_dig[j] = 0xffffffffu;  // not staged (padding or null key)
      t[j] = 0;
      if (r >= tile_n) continue;
      if (FIRST) {
        const unsigned long long vbits = load_value_bits(in_vals, vw, a.vsigned, r);
        const bool vv = (vvalid >> j) & 1u;
        if (!((kvalid >> j) & 1u)) {  // null key: straight into the null group's accumulator
          ++null_rows;
          if (vv) {
            null_sum += vbits;
            ++null_cnt;
          }
          continue;
        }
        const unsigned long long kp = (load_key_bits(in_keys, kw, r) ^ enc.kflip) - enc.kmin;
        const unsigned long long vp = vbits - enc.vbase;
        if (vv && (vp & ~vmask)) atomicOr(a.overflow, 1u);
        t[j] = (kp << (enc.vb + 1)) | (vv ? ((vp & vmask) << 1) | 1ull : 0ull);
      } else {
        t[j] = in_tuples[r];
      }
      const uint32_t d = (compact_hash(compact_key(t[j], enc) + enc.kmin) >> a.shift) & (kPartRadix - 1);
      rank_dig[j] = atomicAdd(&s_cnt[d], 1u) | (d << 16);
    }
    __syncthreads();  // every row of the tile is in registers and counted: the input buffer is free

    if (tid == 0) {
      s_cur = ticket_ahead;
      ticket_ahead = ticket_ahead < n_tiles ? atomicAdd(a.ticket, 1u) : ticket_ahead;  // consumed one iteration from now
    }
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
    const uint32_t next_tile = s_cur;
    if (next_tile < n_tiles) load_tile(next_tile);  // the copy engine fills the buffer while this tile is staged and written
    uint32_t bin_off = 0;
    if (tid < kPartRadix) {
      uint32_t woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < kPartRadix / 32; ++w) {
        if (w < (int)warp) woff += s_warp_tot[w];
        tot += s_warp_tot[w];
      }
      bin_off = woff + incl - run;
      s_bin[tid] = bin_off;
      s_cnt[tid] = 0;  // for the next tile
      if (tid == 0) s_total = tot;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kCItems; ++j) {
      if (rank_dig[j] == 0xffffffffu) continue;
      const uint32_t pos = s_bin[rank_dig[j] >> 16] + (rank_dig[j] & 0xffffu);
      stage[pos] = t[j];
      s_dig[pos] = static_cast<uint8_t>(rank_dig[j] >> 16);
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
    const uint32_t total = s_total;
#pragma unroll
    for (int j = 0; j < kCItems; ++j) {
      const uint32_t p = j * kCThreads + tid;
      if (p < total) {
        __stcs(a.out + s_gbase[s_dig[p]] + p, stage[p]);
      }
    }
    // no barrier here: the next tile writes `stage`, `s_dig` and `s_gbase` only after its own two barriers (counts complete,
    // offsets ready), which no thread passes before it has finished this write-out
    tile = next_tile;
  }

  if (FIRST) {
    // block-reduce the null-key accumulators (rare: usually all zero)
    null_sum = static_cast<unsigned long long>(block_sum<kCThreads>(static_cast<int64_t>(null_sum)));
    __syncthreads();
    null_cnt = static_cast<unsigned long long>(block_sum<kCThreads>(static_cast<int64_t>(null_cnt)));
    __syncthreads();
    null_rows = static_cast<unsigned long long>(block_sum<kCThreads>(static_cast<int64_t>(null_rows)));
    if (tid == 0 && null_rows) {
      atomicAdd(&a.null_acc[0], null_sum);
      atomicAdd(&a.null_acc[1], null_cnt);
      atomicAdd(&a.null_acc[2], null_rows);
    }
  }
}

// ---- pre-aggregation of the partitioned tuples in shared memory, flush to the global table ----------
constexpr int kCChunk = 4096;       // tuples per bulk copy (32 KB), double-buffered
constexpr int kCSliceChunks = 2;    // the shared table is flushed every 8192 tuples
constexpr int kCSlots = 2048;
constexpr int kCProbe = 32;

template <bool NARROW>
constexpr size_t compact_preagg_smem() {
  return 2 * (size_t)kCChunk * 8 + (NARROW ? (size_t)kCSlots * 12 : (size_t)kCSlots * 20);
}

// NARROW: kb <= 31 and vb <= 19 -> 32-bit keys (empty = 0xffffffff), 32-bit partial sums (8192 x 2^19 < 2^32), 32-bit counts
template <bool NARROW>
__global__ void __launch_bounds__(kCThreads, 2) compact_preagg_kernel(const unsigned long long* __restrict__ tuples, uint32_t n,
                                                                      CompactEnc enc, FusedTableRef table,
                                                                      unsigned long long* counters) {
  extern __shared__ __align__(128) uint8_t smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem);  // [2][kCChunk]
  uint8_t* tab = smem + 2 * (size_t)kCChunk * 8;
  uint32_t* k32 = reinterpret_cast<uint32_t*>(tab);                        // NARROW: [slots] keys, [slots] sums, [slots] counts
  uint32_t* s32 = k32 + kCSlots;
  uint32_t* c32 = s32 + kCSlots;
  unsigned long long* k64 = reinterpret_cast<unsigned long long*>(tab);    // WIDE: [slots] keys, [slots] sums (two halves), [slots] counts
  unsigned long long* s64 = k64 + kCSlots;
  uint32_t* c64 = reinterpret_cast<uint32_t*>(s64 + kCSlots);
  __shared__ __align__(8) uint64_t s_bar[2];

  const unsigned tid = threadIdx.x;
  const uint32_t n_chunks = (n + kCChunk - 1) / kCChunk;
  const uint32_t n_slices = (n_chunks + kCSliceChunks - 1) / kCSliceChunks;
  const unsigned long long vmask = (1ull << enc.vb) - 1ull;
  // this CTA's chunk sequence: slices blockIdx.x, blockIdx.x + grid, ... each kCSliceChunks consecutive chunks
  auto chunk_of = [&](uint32_t c) { return (blockIdx.x + (c / kCSliceChunks) * gridDim.x) * kCSliceChunks + (c % kCSliceChunks); };
  uint32_t my_slices = blockIdx.x < n_slices ? (n_slices - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t my_chunks = my_slices * kCSliceChunks;
  auto issue = [&](uint32_t c) {  // thread 0 only
    const uint32_t g = chunk_of(c);
    if (g >= n_chunks) {
      mbar_arrive_expect_tx(&s_bar[c & 1], 0);
      return;
    }
    const uint32_t rows = (n - g * kCChunk) < (uint32_t)kCChunk ? (n - g * kCChunk) : (uint32_t)kCChunk;
    const uint32_t bytes = ((rows * 8u) + 15u) & ~15u;
    mbar_arrive_expect_tx(&s_bar[c & 1], bytes);
    bulk_copy_g2s_chunked(buf + (size_t)(c & 1) * kCChunk, tuples + (size_t)g * kCChunk, bytes, &s_bar[c & 1]);
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
    if (my_chunks > 0) issue(0);
  }
  __syncthreads();
  uint32_t parity[2] = {0, 0};
  for (uint32_t c = 0; c < my_chunks; ++c) {
    if (tid == 0 && c + 1 < my_chunks) issue(c + 1);  // buffer (c+1)&1 was released by the barrier that ended chunk c-1
    if (c % kCSliceChunks == 0) {
      for (int i = tid; i < kCSlots; i += kCThreads) {
        if (NARROW) {
          k32[i] = 0xffffffffu;
          s32[i] = 0;
          c32[i] = 0;
        } else {
          k64[i] = kEmptyKey;
          s64[i] = 0;
          c64[i] = 0;
        }
      }
      __syncthreads();
    }
    mbar_wait(&s_bar[c & 1], parity[c & 1]);
    parity[c & 1] ^= 1u;
    const uint32_t g = chunk_of(c);
    const uint32_t rows = g >= n_chunks ? 0u : ((n - g * kCChunk) < (uint32_t)kCChunk ? (n - g * kCChunk) : (uint32_t)kCChunk);
    const unsigned long long* src = buf + (size_t)(c & 1) * kCChunk;
    // rows are handled kB at a time per thread: all tuples, hashes and first probes of a batch are issued before the
    // dependent shared-memory work of any of them (the serial version ran at 28 % issue utilisation, stalled on LDS -> ATOMS chains)
    constexpr int kB = 4;
    for (uint32_t i0 = tid; i0 < rows; i0 += kB * kCThreads) {
      unsigned long long tt[kB];
      unsigned s0[kB];
      unsigned long long first[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const uint32_t i = i0 + u * kCThreads;
        tt[u] = i < rows ? src[i] : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        s0[u] = compact_hash((tt[u] >> (enc.vb + 1)) + enc.kmin) & (kCSlots - 1);  // low bits: independent of the two partition digits
        first[u] = NARROW ? static_cast<unsigned long long>(k32[s0[u]]) : k64[s0[u]];
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (i0 + u * kCThreads >= rows) continue;
        const unsigned long long t = tt[u];
        const unsigned long long kp = t >> (enc.vb + 1);
        const unsigned long long vp = (t >> 1) & vmask;
        const bool vv = t & 1ull;
        unsigned s = s0[u];
        int slot = -1;
        for (int probe = 0; probe < kCProbe; ++probe) {
          if (NARROW) {
            const uint32_t cur = probe == 0 ? static_cast<uint32_t>(first[u]) : k32[s];
            if (cur == (uint32_t)kp) { slot = s; break; }
            if (cur == 0xffffffffu) {
              const uint32_t old = atomicCAS(&k32[s], 0xffffffffu, (uint32_t)kp);
              if (old == 0xffffffffu || old == (uint32_t)kp) { slot = s; break; }
            }
          } else {
            const unsigned long long cur = probe == 0 ? first[u] : k64[s];
            if (cur == kp) { slot = s; break; }
            if (cur == kEmptyKey) {
              const unsigned long long old = atomicCAS(&k64[s], (unsigned long long)kEmptyKey, kp);
              if (old == kEmptyKey || old == kp) { slot = s; break; }
            }
          }
          s = (s + 1) & (kCSlots - 1);
        }
        if (slot < 0) {  // the slice holds too many distinct keys: this row goes straight to the global table
          global_accumulate<false>(table, (kp + enc.kmin) ^ enc.kflip, false, vp + enc.vbase, vv ? 1u : 0u, counters);
          continue;
        }
        if (NARROW) {
          if (vv) {
            atomicAdd(&s32[slot], (uint32_t)vp);
            atomicAdd(&c32[slot], 1u);
          }
        } else if (vv) {
          unsigned int* half = reinterpret_cast<unsigned int*>(&s64[slot]);  // [0] = lo, [1] = hi; explicit carry (native ATOMS only)
          const unsigned int lo32 = static_cast<unsigned int>(vp), hi32 = static_cast<unsigned int>(vp >> 32);
          const unsigned int old = atomicAdd(half, lo32);
          const unsigned int carry = (old + lo32) < old ? 1u : 0u;
          if (hi32 + carry) atomicAdd(half + 1, hi32 + carry);
          atomicAdd(&c64[slot], 1u);
        }
      }
    }
    __syncthreads();  // chunk consumed: its buffer may be refilled; the table is complete if the slice ends here
    if (c % kCSliceChunks == kCSliceChunks - 1) {
      for (int i = tid; i < kCSlots; i += kCThreads) {
        unsigned long long kp, sum;
        unsigned cnt;
        if (NARROW) {
          if (k32[i] == 0xffffffffu) continue;
          kp = k32[i];
          sum = s32[i];
          cnt = c32[i];
        } else {
          if (k64[i] == kEmptyKey) continue;
          kp = k64[i];
          sum = s64[i];
          cnt = c64[i];
        }
        // sum of the real values = sum of the offsets + count x base (mod 2^64, the reference's wrap-around)
        global_accumulate<false>(table, (kp + enc.kmin) ^ enc.kflip, false, sum + (unsigned long long)cnt * enc.vbase, cnt, counters);
      }
      __syncthreads();
    }
  }
}

// the null group's partial state from pass 1 -> the global table (one thread)
__global__ void compact_null_flush_kernel(FusedTableRef table, const unsigned long long* null_acc, unsigned long long* counters) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && null_acc[2] != 0)
    global_accumulate<false>(table, 0ull, true, null_acc[0], static_cast<unsigned>(null_acc[1] > 0xffffffffull ? 0xffffffffu : null_acc[1]), counters);
}

}  // namespace b2
