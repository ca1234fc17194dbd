// tma.cuh -- bulk asynchronous global -> shared copies (the TMA engine's 1-D mode) + mbarrier.
//
// Used by the partition / pre-aggregation passes (groupby_compact.cuh) to keep the NEXT tile's bytes in
// flight while the current tile is ranked, staged and written: one elected thread arms an mbarrier with
// the byte count and issues cp.async.bulk; every consumer thread waits on the barrier's phase parity.
// SASS: UBLKCP (bulk copy) + SYNCS (mbarrier); no registers are spent on the data in flight.
// Constraints (PTX ISA cp.async.bulk): global address, shared address and size are multiples of 16 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// make the initialised barrier visible to the async proxy (the copy engine) before the first copy is issued
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// one bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// issue a (possibly large) contiguous copy as <= 16 KB pieces; the caller has already armed `bar`
// with the total byte count
__device__ __forceinline__ void bulk_copy_g2s_chunked(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  constexpr uint32_t kPiece = 16384;
  for (uint32_t off = 0; off < bytes; off += kPiece) {
    const uint32_t b = bytes - off < kPiece ? bytes - off : kPiece;
    bulk_copy_g2s(static_cast<char*>(smem_dst) + off, static_cast<const char*>(gmem_src) + off, b, bar);
  }
}

}  // namespace b2
