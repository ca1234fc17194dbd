// smem_probe.cu -- throughput of the shared-memory / warp primitives the radix-partition and
// pre-aggregation kernels lean on (per-SM rates decide whether those kernels are HBM- or MIO-bound).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_probe smem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512) probe(uint64_t* sink, int iters, int spread) {
  __shared__ unsigned long long s64[4096];
  __shared__ unsigned int s32[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) { s64[i] = 0; s32[i] = 0; }
  __syncthreads();
  uint32_t x = mix32(blockIdx.x * 512 + threadIdx.x + 1);
  uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    x = mix32(x + it);
    const unsigned a = x & (spread - 1);
    if (MODE == 0) atomicAdd(&s32[a], 1u);                                   // ATOMS.ADD u32
    else if (MODE == 1) atomicAdd(&s64[a], (unsigned long long)x);           // 64-bit add
    else if (MODE == 2) acc += atomicCAS(&s64[a], 0ull, (unsigned long long)x | 1ull);  // 64-bit CAS
    else if (MODE == 3) acc += __match_any_sync(0xffffffffu, a & 255);       // MATCH.ANY 32-bit
    else if (MODE == 4) acc += __match_any_sync(0xffffffffu, (unsigned long long)a * 0x9E3779B97F4A7C15ull);  // 64-bit
    else if (MODE == 5) acc += atomicAdd(&s32[a], 1u);                       // returning ATOMS
    else if (MODE == 6) { s32[a] += 1; acc += s32[(a + 7) & 4095]; }         // plain LDS/STS (racy, rate only)
    else if (MODE == 7) atomicAdd(reinterpret_cast<double*>(&s64[a]), 1.0);  // double add (CAS loop?)
    else if (MODE == 8) acc += __reduce_add_sync(0xffffffffu, a);            // REDUX
    else if (MODE == 9) acc += __ballot_sync(0xffffffffu, a & 1);            // VOTE
  }
  if (acc == 0x1234567) sink[0] = acc + s64[1] + s32[2];
}

template <int MODE>
void run(const char* name, uint64_t* sink, int spread) {
  const int blocks = 148 * 2, iters = 4096;
  probe<MODE><<<blocks, 512>>>(sink, 64, spread);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<MODE><<<blocks, 512>>>(sink, iters, spread);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * 512 * iters;
  printf("%-34s spread=%5d  %8.3f ms  %8.1f G lane-ops/s  (%.2f cyc/lane/SM @1.9GHz)\n", name, spread, ms, ops / ms / 1e6,
         148 * 1.9e9 / (ops / (ms * 1e-3)));
}

int main() {
  uint64_t* sink; cudaMalloc(&sink, 64);
  for (int spread : {4096, 256, 1}) {
    run<0>("ATOMS.ADD u32 (no return)", sink, spread);
    run<5>("ATOMS.ADD u32 (returning)", sink, spread);
    run<1>("atomicAdd u64 smem", sink, spread);
    run<2>("atomicCAS u64 smem", sink, spread);
    run<7>("atomicAdd double smem", sink, spread);
  }
  run<3>("match_any 32-bit (256 values)", sink, 4096);
  run<4>("match_any 64-bit", sink, 4096);
  run<6>("plain LDS+STS", sink, 4096);
  run<8>("redux.add", sink, 4096);
  run<9>("ballot", sink, 4096);
  printf("# %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
