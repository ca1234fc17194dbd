// elementwise.cuh -- the streaming applicator shared by cast / arithmetic / compare.
//
// Plays the role of the reference's ScalarUnary / ScalarBinary applicators
// (cpp/src/arrow/compute/kernels/codegen_internal.h:590-976): run an op functor over
// every slot (including slots under nulls, like the reference) of one or two inputs.
//
// B200 mapping: a warp owns 32*V*U consecutive elements; each lane issues U
// independent 16-byte (or narrower, for the narrower side of a widening op) streaming
// accesses striped by lane, so every instruction touches one contiguous 512-byte span
// (fully coalesced 128 B lines) and U requests are in flight per lane before the first
// use.  No shared memory: each byte is touched once (guide: element-wise kernels).
#pragma once
#include "common.cuh"

namespace b2 {

template <int NB>
struct Bytes;
template <>
struct Bytes<1> { using type = uint8_t; };
template <>
struct Bytes<2> { using type = uint16_t; };
template <>
struct Bytes<4> { using type = uint32_t; };
template <>
struct Bytes<8> { using type = uint2; };
template <>
struct Bytes<16> { using type = uint4; };

template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> load_vec(const T* p) {
  using B = typename Bytes<sizeof(T) * V>::type;
  B raw = __ldcs(reinterpret_cast<const B*>(p));
  Vec<T, V> r;
  *reinterpret_cast<B*>(&r) = raw;
  return r;
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const Vec<T, V>& x) {
  using B = typename Bytes<sizeof(T) * V>::type;
  __stcs(reinterpret_cast<B*>(p), *reinterpret_cast<const B*>(&x));
}

template <typename A, typename B>
constexpr int vec_elems() {
  return 16 / (sizeof(A) > sizeof(B) ? sizeof(A) : sizeof(B));
}

constexpr int kUnroll = 4;

// out[i] = f(in[i], i)
template <typename In, typename Out, int V, typename F>
__global__ void __launch_bounds__(kBlock) map1_kernel(const In* __restrict__ in,
                                                      Out* __restrict__ out, int64_t n,
                                                      bool vec_ok, F f) {
  constexpr int U = kUnroll;
  constexpr int64_t kWarpTile = 32 * V * U;
  constexpr int64_t kTile = kWarpTile * kWarpsPerBlock;
  const unsigned lane = lane_id();
  for (int64_t tile = (int64_t)blockIdx.x * kTile; tile < n; tile += (int64_t)gridDim.x * kTile) {
    int64_t wb = tile + (int64_t)(threadIdx.x >> 5) * kWarpTile;
    if (wb >= n) continue;
    if (vec_ok && wb + kWarpTile <= n) {
      Vec<In, V> x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = load_vec<In, V>(in + wb + u * 32 * V + lane * V);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        Vec<Out, V> y;
        int64_t i0 = wb + u * 32 * V + lane * V;
#pragma unroll
        for (int k = 0; k < V; ++k) y.v[k] = f(x[u].v[k], i0 + k);
        store_vec<Out, V>(out + i0, y);
      }
    } else {
      int64_t end = wb + kWarpTile < n ? wb + kWarpTile : n;
      for (int64_t i = wb + lane; i < end; i += 32) out[i] = f(in[i], i);
    }
  }
}

// out[i] = f(a[i] or sa, b[i] or sb, i); a == nullptr means broadcast scalar sa
template <typename In, typename Out, int V, typename F>
__global__ void __launch_bounds__(kBlock) map2_kernel(const In* __restrict__ a, In sa,
                                                      const In* __restrict__ b, In sb,
                                                      Out* __restrict__ out, int64_t n,
                                                      bool vec_ok, F f) {
  constexpr int U = kUnroll;
  constexpr int64_t kWarpTile = 32 * V * U;
  constexpr int64_t kTile = kWarpTile * kWarpsPerBlock;
  const unsigned lane = lane_id();
  for (int64_t tile = (int64_t)blockIdx.x * kTile; tile < n; tile += (int64_t)gridDim.x * kTile) {
    int64_t wb = tile + (int64_t)(threadIdx.x >> 5) * kWarpTile;
    if (wb >= n) continue;
    if (vec_ok && wb + kWarpTile <= n) {
      Vec<In, V> x[U], y[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t i0 = wb + u * 32 * V + lane * V;
        if (a) {
          x[u] = load_vec<In, V>(a + i0);
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) x[u].v[k] = sa;
        }
        if (b) {
          y[u] = load_vec<In, V>(b + i0);
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) y[u].v[k] = sb;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        Vec<Out, V> z;
        int64_t i0 = wb + u * 32 * V + lane * V;
#pragma unroll
        for (int k = 0; k < V; ++k) z.v[k] = f(x[u].v[k], y[u].v[k], i0 + k);
        store_vec<Out, V>(out + i0, z);
      }
    } else {
      int64_t end = wb + kWarpTile < n ? wb + kWarpTile : n;
      for (int64_t i = wb + lane; i < end; i += 32) out[i] = f(a ? a[i] : sa, b ? b[i] : sb, i);
    }
  }
}

template <typename In, typename Out, int V>
inline int map_grid(int64_t n) {
  constexpr int64_t kTile = (int64_t)32 * V * kUnroll * kWarpsPerBlock;
  // up to 16 CTA-waves over the 148 SMs; beyond that CTAs loop (grid-stride)
  return grid_for(n, kTile, kSMs * 8 * 16);
}

// first-error cell shared by the checked ops: rows race with atomicMin so the
// reported element is the first offending row, as the reference's in-order scan finds.
struct ErrorCell {
  unsigned long long* first_row;  // device; initialised to ~0ull
  __device__ __forceinline__ void report(int64_t row) const {
    atomicMin(first_row, static_cast<unsigned long long>(row));
  }
};

}  // namespace b2
