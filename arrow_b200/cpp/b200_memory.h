// b200_memory.h -- device memory for the arrow::compute drop-in: an arrow::Device /
// arrow::MemoryManager / arrow::Buffer triple backed by the libarrow_b200 pool.
//
// B200-native counterpart of arrow::cuda::CudaDevice / CudaMemoryManager / CudaBuffer
// (cpp/src/arrow/gpu/cuda_context.h:253, cuda_memory.h:39-110; cuda_context.cc:110-121
// allocates with one cuMemAlloc per buffer).  Buffers report device_type() == kCUDA, are
// not CPU-accessible (Buffer::data() == nullptr, buffer.h:225) and expose their device
// address through Buffer::address() (buffer.h:277) -- which is what the kernel trampolines
// read via ArraySpan.buffers[i].owner (array/data.h:525-532).
#pragma once
#include <arrow/api.h>
#include <arrow/c/abi.h>
#include <arrow/device.h>

#include "arrow_b200.h"

namespace arrow_b200 {

class B200MemoryManager;

class B200Device : public arrow::Device {
 public:
  static arrow::Result<std::shared_ptr<B200Device>> Make(int device_number = 0);
  ~B200Device() override;
  const char* type_name() const override { return "arrow_b200::B200Device"; }
  std::string ToString() const override;
  bool Equals(const arrow::Device& other) const override;
  int64_t device_id() const override { return device_number_; }
  std::shared_ptr<arrow::MemoryManager> default_memory_manager() override;
  arrow::DeviceAllocationType device_type() const override { return arrow::DeviceAllocationType::kCUDA; }
  B2Context* context() const { return ctx_; }
  // the context's own stream as an arrow::Device::Stream (SyncEvent::Record / Stream::WaitEvent operate on it)
  arrow::Result<std::shared_ptr<arrow::Device::Stream>> MakeStream() override;
  arrow::Result<std::shared_ptr<arrow::Device::Stream>> WrapStream(void* cuda_stream_ptr,
                                                                   arrow::Device::Stream::release_fn_t release) override;

 private:
  explicit B200Device(int n, B2Context* ctx) : arrow::Device(/*is_cpu=*/false), device_number_(n), ctx_(ctx) {}
  int device_number_;
  B2Context* ctx_;
  std::weak_ptr<arrow::MemoryManager> mm_;
};

class B200MemoryManager : public arrow::MemoryManager {
 public:
  explicit B200MemoryManager(const std::shared_ptr<arrow::Device>& device) : arrow::MemoryManager(device) {}
  arrow::Result<std::shared_ptr<arrow::io::RandomAccessFile>> GetBufferReader(std::shared_ptr<arrow::Buffer> buf) override;
  arrow::Result<std::shared_ptr<arrow::io::OutputStream>> GetBufferWriter(std::shared_ptr<arrow::Buffer> buf) override;
  arrow::Result<std::unique_ptr<arrow::Buffer>> AllocateBuffer(int64_t size) override;
  B2Context* context() const { return static_cast<B200Device*>(device_.get())->context(); }
  // take ownership of a pool pointer returned by a C-ABI entry point
  std::shared_ptr<arrow::Buffer> Adopt(const void* ptr, int64_t size);
  // C Device Data Interface synchronisation: a cudaEvent_t behind arrow::Device::SyncEvent (device.h:142-165);
  // get_raw() is the cudaEvent_t* that ArrowDeviceArray.sync_event carries for ARROW_DEVICE_CUDA
  arrow::Result<std::shared_ptr<arrow::Device::SyncEvent>> MakeDeviceSyncEvent() override;
  arrow::Result<std::shared_ptr<arrow::Device::SyncEvent>> WrapDeviceSyncEvent(
      void* sync_event, arrow::Device::SyncEvent::release_fn_t release_sync_event) override;

 protected:
  arrow::Result<std::shared_ptr<arrow::Buffer>> CopyBufferFrom(const std::shared_ptr<arrow::Buffer>& buf,
                                                               const std::shared_ptr<arrow::MemoryManager>& from) override;
  arrow::Result<std::shared_ptr<arrow::Buffer>> CopyBufferTo(const std::shared_ptr<arrow::Buffer>& buf,
                                                             const std::shared_ptr<arrow::MemoryManager>& to) override;
  arrow::Result<std::unique_ptr<arrow::Buffer>> CopyNonOwnedFrom(const arrow::Buffer& buf,
                                                                 const std::shared_ptr<arrow::MemoryManager>& from) override;
  arrow::Result<std::unique_ptr<arrow::Buffer>> CopyNonOwnedTo(const arrow::Buffer& buf,
                                                               const std::shared_ptr<arrow::MemoryManager>& to) override;
};

// Status built from the calling thread's b2_last_error() with the matching StatusCode
arrow::Status StatusFromB2(int code);
#define B200_RETURN_NOT_OK(expr)                              \
  do {                                                        \
    int _b2s = (expr);                                        \
    if (_b2s != 0) return ::arrow_b200::StatusFromB2(_b2s);   \
  } while (0)

// whole-array transfers (Array::CopyTo works too; these keep null_count and offset)
arrow::Result<std::shared_ptr<arrow::ArrayData>> ToDevice(const arrow::ArrayData& host,
                                                          const std::shared_ptr<arrow::MemoryManager>& mm);
arrow::Result<std::shared_ptr<arrow::ArrayData>> ToHost(const arrow::ArrayData& device);
bool IsOnDevice(const arrow::ArrayData& data);

// ---- C Device Data Interface (c/bridge.h:192,238) for pool-backed arrays ----
// Export: records a fresh event on the context stream (every kernel that produced the buffers was ordered
// on it) and hands the array out as an ArrowDeviceArray{device_type = ARROW_DEVICE_CUDA, sync_event = cudaEvent_t*};
// the exported structure keeps the buffers and the event alive until its release callback runs.
arrow::Status ExportDeviceArray(const arrow::Array& array, const std::shared_ptr<arrow::MemoryManager>& mm,
                                struct ArrowDeviceArray* out, struct ArrowSchema* out_schema = nullptr);
// Import: wraps the producer's device pointers (zero copy) in buffers of OUR memory manager and makes the
// context stream wait on the producer's sync_event before any kernel of ours can touch them.
arrow::Result<std::shared_ptr<arrow::Array>> ImportDeviceArray(struct ArrowDeviceArray* array, std::shared_ptr<arrow::DataType> type,
                                                               const std::shared_ptr<arrow::MemoryManager>& mm);

// ---- C Device STREAM interface (c/abi.h ArrowDeviceArrayStream, c/bridge.h:334,386): ingest of a producer's stream of
// device record batches.  Every batch is wrapped zero-copy in buffers of OUR memory manager (the DeviceMemoryMapper hands
// it out for ARROW_DEVICE_CUDA) and the context stream is ordered behind the producer's per-buffer sync events before
// the batch is returned, so the batches can go straight into the kernels / a `record_batch_reader_source` of a b200_ plan.
arrow::Result<std::shared_ptr<arrow::RecordBatchReader>> ImportDeviceRecordBatchReader(struct ArrowDeviceArrayStream* stream,
                                                                                     const std::shared_ptr<arrow::MemoryManager>& mm);

}  // namespace arrow_b200