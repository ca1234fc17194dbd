// context.cu -- B2Context: device pool, streams, scalar read-back, error strings.
//
// The reference's device memory path is arrow::cuda::CudaContext::Allocate
// (cpp/src/arrow/gpu/cuda_context.cc:110-121: one cuMemAlloc per buffer, no pool) and
// CudaHostBuffer for pinned staging (cpp/src/arrow/gpu/cuda_memory.h:113).  Kernels
// here allocate an output and a few temporaries per call, so the B200 build keeps a
// caching pool: freed blocks are binned by size and handed back stream-ordered,
// cudaMalloc/cudaFree only happen on first touch or under memory pressure.
#include "context.h"

#include <algorithm>
#include <cstring>

namespace b2 {

std::atomic<int64_t> g_launches{0};
static thread_local std::string t_last_error;

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

static size_t round_size(size_t n) {
  if (n <= 256) return 256;
  if (n <= (1u << 20)) {  // power of two below 1 MiB
    size_t r = 256;
    while (r < n) r <<= 1;
    return r;
  }
  const size_t g = 2u << 20;  // 2 MiB granules above (the TLB page size)
  return (n + g - 1) / g * g;
}

ScalarSlot::ScalarSlot(B2Context* c) : ctx(c) {
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->free_slots.empty()) {
    idx = c->free_slots.back();
    c->free_slots.pop_back();
  }
}
ScalarSlot::~ScalarSlot() {
  if (idx >= 0) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->free_slots.push_back(idx);
  }
}
int ScalarSlot::zero(cudaStream_t s) {
  if (idx < 0) return set_error(B2_UNKNOWN_ERROR, "no free scalar slot");
  B2_CUDA(cudaMemsetAsync(dev(), 0, B2Context::kSlotBytes, s));
  return B2_OK;
}
int ScalarSlot::fetch(cudaStream_t s) {
  B2_CUDA(cudaMemcpyAsync(const_cast<int64_t*>(host()), dev(), B2Context::kSlotBytes,
                          cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  return B2_OK;
}

}  // namespace b2

using b2::set_error;

int B2Context::alloc(size_t nbytes, void** out, cudaStream_t s) {
  *out = nullptr;
  if (nbytes == 0) nbytes = 1;
  if (user_alloc) {
    void* p = user_alloc(nbytes, s, user_data);
    if (!p) return set_error(B2_OUT_OF_MEMORY, "allocator callback failed for %zu bytes", nbytes);
    std::lock_guard<std::mutex> lk(mu);
    user_sizes[p] = nbytes;
    *out = p;
    return B2_OK;
  }
  size_t want = b2::round_size(nbytes);
  cudaStream_t wait_on = nullptr;
  bool need_wait = false;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = free_blocks.lower_bound(want);
    // accept a cached block up to 25% larger than requested
    if (it != free_blocks.end() && it->first <= want + want / 4) {
      void* p = it->second;
      size_t sz = it->first;
      free_blocks.erase(it);
      Block& b = blocks[p];
      if (b.last_stream != s) {
        need_wait = true;
        wait_on = b.last_stream;
      }
      b.last_stream = s;
      in_use[p] = sz;
      bytes_in_use += static_cast<int64_t>(sz);
      max_in_use = std::max(max_in_use, bytes_in_use);
      *out = p;
    }
  }
  if (*out) {
    // a block last used on another stream may still be read by work queued there
    if (need_wait) B2_CUDA(cudaStreamSynchronize(wait_on));
    return B2_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    trim();
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(B2_OUT_OF_MEMORY, "cudaMalloc of %zu bytes failed: %s", want,
                     cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(mu);
  blocks[p] = Block{want, s};
  in_use[p] = want;
  bytes_in_use += static_cast<int64_t>(want);
  bytes_reserved += static_cast<int64_t>(want);
  max_in_use = std::max(max_in_use, bytes_in_use);
  *out = p;
  return B2_OK;
}

int B2Context::free(void* ptr, cudaStream_t s) {
  if (!ptr) return B2_OK;
  if (user_free) {
    size_t sz = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = user_sizes.find(ptr);
      if (it != user_sizes.end()) {
        sz = it->second;
        user_sizes.erase(it);
      }
    }
    user_free(ptr, sz, s, user_data);
    return B2_OK;
  }
  std::lock_guard<std::mutex> lk(mu);
  auto it = in_use.find(ptr);
  if (it == in_use.end()) return set_error(B2_INVALID, "b2_free: pointer %p not from this pool", ptr);
  size_t sz = it->second;
  in_use.erase(it);
  bytes_in_use -= static_cast<int64_t>(sz);
  blocks[ptr].last_stream = s;
  free_blocks.emplace(sz, ptr);
  return B2_OK;
}

int B2Context::trim() {
  std::vector<void*> victims;
  {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : free_blocks) {
      victims.push_back(kv.second);
      bytes_reserved -= static_cast<int64_t>(kv.first);
      blocks.erase(kv.second);
    }
    free_blocks.clear();
  }
  if (!victims.empty()) cudaDeviceSynchronize();
  for (void* p : victims) cudaFree(p);
  return B2_OK;
}

extern "C" {

int b2_context_create(int device, B2Context** out) {
  if (!out) return set_error(B2_INVALID, "b2_context_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(B2_CUDA_ERROR,
                     "arrow_b200 needs a CUDA device; none is visible (%s). There is no CPU "
                     "fallback.",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return set_error(B2_INVALID, "device %d out of range [0,%d)", device, n);
  B2_CUDA(cudaSetDevice(device));
  B2Context* c = new B2Context();
  c->device = device;
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  B2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  B2_CUDA(cudaMalloc(&c->slot_dev, B2Context::kSlots * B2Context::kSlotBytes));
  B2_CUDA(cudaHostAlloc(&c->slot_host, B2Context::kSlots * B2Context::kSlotBytes,
                        cudaHostAllocDefault));
  for (int i = 0; i < B2Context::kSlots; ++i) c->free_slots.push_back(i);
  *out = c;
  return B2_OK;
}

void b2_context_destroy(B2Context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->trim();
  for (auto& kv : ctx->in_use) cudaFree(kv.first);
  if (ctx->slot_dev) cudaFree(ctx->slot_dev);
  if (ctx->slot_host) cudaFreeHost(ctx->slot_host);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int b2_context_device(const B2Context* ctx) { return ctx ? ctx->device : -1; }
void* b2_context_stream(const B2Context* ctx) { return ctx ? ctx->stream : nullptr; }

int b2_context_set_allocator(B2Context* ctx, B2AllocFn alloc, B2FreeFn free_fn, void* user) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  if ((alloc == nullptr) != (free_fn == nullptr))
    return set_error(B2_INVALID, "alloc and free callbacks must be set together");
  ctx->user_alloc = alloc;
  ctx->user_free = free_fn;
  ctx->user_data = user;
  return B2_OK;
}

int b2_alloc(B2Context* ctx, size_t nbytes, void** out) {
  if (!ctx || !out) return set_error(B2_INVALID, "b2_alloc: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  return ctx->alloc(nbytes, out, ctx->stream);
}
int b2_free(B2Context* ctx, void* ptr) {
  if (!ctx) return set_error(B2_INVALID, "b2_free: null context");
  return ctx->free(ptr, ctx->stream);
}
int b2_pool_stats(const B2Context* ctx, int64_t* in_use, int64_t* reserved, int64_t* max_in_use) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  if (in_use) *in_use = ctx->bytes_in_use;
  if (reserved) *reserved = ctx->bytes_reserved;
  if (max_in_use) *max_in_use = ctx->max_in_use;
  return B2_OK;
}
int b2_pool_trim(B2Context* ctx) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  return ctx->trim();
}
int b2_sync(B2Context* ctx, void* stream) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  B2_CUDA(cudaStreamSynchronize(ctx->pick(stream)));
  return B2_OK;
}
int b2_memcpy_h2d(B2Context* ctx, void* dst, const void* src, size_t n, void* stream) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  if (n == 0) return B2_OK;
  B2_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->pick(stream)));
  return B2_OK;
}
int b2_memcpy_d2h(B2Context* ctx, void* dst, const void* src, size_t n, void* stream) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  if (n == 0) return B2_OK;
  B2_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->pick(stream)));
  return B2_OK;
}
int b2_memset(B2Context* ctx, void* dst, int byte, size_t n, void* stream) {
  if (!ctx) return set_error(B2_INVALID, "null context");
  if (n == 0) return B2_OK;
  B2_CUDA(cudaMemsetAsync(dst, byte, n, ctx->pick(stream)));
  return B2_OK;
}
int b2_host_alloc(size_t nbytes, void** out) {
  if (!out) return set_error(B2_INVALID, "null out");
  B2_CUDA(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
  return B2_OK;
}
int b2_host_free(void* ptr) {
  if (ptr) B2_CUDA(cudaFreeHost(ptr));
  return B2_OK;
}

// ---- events: what ArrowDeviceArray.sync_event points at for ARROW_DEVICE_CUDA is a cudaEvent_t ----
int b2_event_create(B2Context* ctx, void** out_event) {
  if (!ctx || !out_event) return set_error(B2_INVALID, "b2_event_create: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaEvent_t e;
  B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out_event = e;
  return B2_OK;
}
int b2_event_destroy(void* event) {
  if (event) B2_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return B2_OK;
}
int b2_event_record(B2Context* ctx, void* event, void* stream) {
  if (!ctx || !event) return set_error(B2_INVALID, "b2_event_record: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), ctx->pick(stream)));
  return B2_OK;
}
int b2_event_synchronize(void* event) {
  if (!event) return set_error(B2_INVALID, "b2_event_synchronize: null argument");
  B2_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return B2_OK;
}
int b2_stream_wait_event(B2Context* ctx, void* stream, void* event) {
  if (!ctx || !event) return set_error(B2_INVALID, "b2_stream_wait_event: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaStreamWaitEvent(ctx->pick(stream), static_cast<cudaEvent_t>(event), 0));
  return B2_OK;
}

const char* b2_last_error(void) { return b2::t_last_error.c_str(); }
const char* b2_version(void) { return "arrow_b200 0.1.0 (sm_100a)"; }
int64_t b2_launch_count(void) { return b2::g_launches.load(); }

}  // extern "C"
