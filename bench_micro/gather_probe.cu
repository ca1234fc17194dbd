// gather_probe.cu -- microbenchmark behind the Take design: random 8-byte gathers per second as
// a function of table size (TLB reach / L2 residency) and of gathers in flight per lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k;
}

template <int U>
__global__ void gather_kernel(const uint64_t* __restrict__ vals, const int64_t* __restrict__ idx, uint64_t* __restrict__ out,
                              int64_t n, uint64_t mask, int64_t window_rows) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * U) {
    int64_t j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + u * stride;
      j[u] = i < n ? (idx ? idx[i] : (int64_t)(mix(i) & mask)) : 0;
      if (window_rows) j[u] = (j[u] % window_rows) + (i / window_rows) % ((mask + 1) / window_rows) * window_rows;
    }
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = vals[j[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) { int64_t i = base + u * stride; if (i < n) out[i] = v[u]; }
  }
}

__global__ void fill_idx(int64_t* idx, int64_t n, uint64_t mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) idx[i] = mix(i * 7 + 1) & mask;
}

template <int U>
float run(const uint64_t* vals, const int64_t* idx, uint64_t* out, int64_t n, uint64_t mask, int64_t window, int blocks) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather_kernel<U><<<blocks, 256>>>(vals, idx, out, n, mask, window);
  cudaEventRecord(a);
  for (int r = 0; r < 3; ++r) gather_kernel<U><<<blocks, 256>>>(vals, idx, out, n, mask, window);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}

int main() {
  const int64_t n = 1ll << 28;  // 268M gathers per launch
  uint64_t *vals, *out; int64_t* idx;
  cudaMalloc(&vals, 8ull << 30); cudaMalloc(&out, n * 8); cudaMalloc(&idx, n * 8);
  cudaMemset(vals, 1, 8ull << 30);
  printf("table_MB,U,idx_in_mem,window_MB,ms,Ggathers_per_s\n");
  for (int shift = 23; shift <= 30; ++shift) {  // table rows 2^shift * 8 B = 64 MB .. 8 GB
    uint64_t mask = (1ull << shift) - 1;
    fill_idx<<<148 * 8, 256>>>(idx, n, mask);
    for (int mem = 0; mem < 2; ++mem) {
      float t4 = run<4>(vals, mem ? idx : nullptr, out, n, mask, 0, 148 * 16);
      float t8 = run<8>(vals, mem ? idx : nullptr, out, n, mask, 0, 148 * 16);
      float t16 = run<16>(vals, mem ? idx : nullptr, out, n, mask, 0, 148 * 16);
      printf("%llu,4,%d,0,%.3f,%.1f\n", (unsigned long long)((mask + 1) * 8 >> 20), mem, t4, n / t4 / 1e6);
      printf("%llu,8,%d,0,%.3f,%.1f\n", (unsigned long long)((mask + 1) * 8 >> 20), mem, t8, n / t8 / 1e6);
      printf("%llu,16,%d,0,%.3f,%.1f\n", (unsigned long long)((mask + 1) * 8 >> 20), mem, t16, n / t16 / 1e6);
    }
  }
  // L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 bytes), 8 GB table
  for (size_t gran : {32, 64, 128}) {
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    uint64_t m = (1ull << 30) - 1;
    float t = run<8>(vals, nullptr, out, n, m, 0, 148 * 16);
    printf("# l2_fetch_granularity req=%zu got=%zu : %.3f ms, %.1f Ggathers/s\n", gran, got, t, n / t / 1e6);
  }
  // 8 GB table, but consecutive groups of gathers confined to a window (what a partition pass would give)
  uint64_t mask = (1ull << 30) - 1;
  for (int wshift = 22; wshift <= 28; wshift += 2) {
    int64_t window = 1ll << wshift;
    float t = run<8>(vals, nullptr, out, n, mask, window, 148 * 16);
    printf("8192,8,0,%lld,%.3f,%.1f\n", (long long)(window * 8 >> 20), t, n / t / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("# %s\n", cudaGetErrorString(e));
  return 0;
}
