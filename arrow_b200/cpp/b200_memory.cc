// b200_memory.cc -- see b200_memory.h
#include "b200_memory.h"

#include <arrow/c/bridge.h>
#include <arrow/io/interfaces.h>
#include <arrow/util/logging.h>

#include <map>
#include <mutex>

namespace arrow_b200 {

using arrow::Result;
using arrow::Status;

Status StatusFromB2(int code) {
  const char* msg = b2_last_error();
  switch (code) {
    case B2_OUT_OF_MEMORY: return Status::OutOfMemory(msg);
    case B2_KEY_ERROR: return Status::KeyError(msg);
    case B2_TYPE_ERROR: return Status::TypeError(msg);
    case B2_INVALID: return Status::Invalid(msg);
    case B2_IO_ERROR: return Status::IOError(msg);
    case B2_CAPACITY_ERROR: return Status::CapacityError(msg);
    case B2_INDEX_ERROR: return Status::IndexError(msg);
    case B2_NOT_IMPLEMENTED: return Status::NotImplemented(msg);
    default: return Status::UnknownError(msg);
  }
}

namespace {

// A pool-owned device buffer: freed back into the libarrow_b200 pool on destruction,
// as CudaBuffer frees through its context (cuda_memory.cc).
class B200Buffer : public arrow::MutableBuffer {
 public:
  B200Buffer(uint8_t* ptr, int64_t size, std::shared_ptr<arrow::MemoryManager> mm, B2Context* ctx)
      : arrow::MutableBuffer(ptr, size, std::move(mm)), ctx_(ctx), ptr_(ptr) {}
  ~B200Buffer() override {
    if (ptr_) b2_free(ctx_, ptr_);
  }

 private:
  B2Context* ctx_;
  void* ptr_;
};

// Result buffers of device->host copies come from a cache of PINNED host blocks (the role of
// CudaHostBuffer / AllocateCudaHostBuffer, gpu/cuda_memory.h:113,254): a pageable destination makes the driver
// stage the copy and first-touches every page (measured ~15 ms for the 240 MB group table of config 3); a
// pinned one runs at the link rate and, once cached, costs nothing to allocate.  They are ordinary CPU buffers to
// every consumer (is_cpu(), default CPU memory manager).
class PinnedCache {
 public:
  static PinnedCache& Get() {
    static PinnedCache* cache = new PinnedCache();  // outlives every buffer handed out
    return *cache;
  }
  Result<void*> Acquire(size_t size, size_t* capacity) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      auto it = free_.lower_bound(size);
      if (it != free_.end() && it->first <= 2 * size + (1 << 20)) {
        void* p = it->second;
        *capacity = it->first;
        cached_ -= it->first;
        free_.erase(it);
        return p;
      }
    }
    const size_t cap = (size + (1 << 20) - 1) & ~static_cast<size_t>((1 << 20) - 1);
    void* p = nullptr;
    B200_RETURN_NOT_OK(b2_host_alloc(cap, &p));
    *capacity = cap;
    return p;
  }
  void Release(void* p, size_t capacity) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      if (cached_ + capacity <= kMaxCached) {
        free_.emplace(capacity, p);
        cached_ += capacity;
        return;
      }
    }
    b2_host_free(p);
  }

 private:
  static constexpr size_t kMaxCached = 4ull << 30;
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
  size_t cached_ = 0;
};

class PinnedResultBuffer : public arrow::MutableBuffer {
 public:
  PinnedResultBuffer(void* ptr, int64_t size, size_t capacity)
      : arrow::MutableBuffer(static_cast<uint8_t*>(ptr), size), ptr_(ptr), capacity_(capacity) {}
  ~PinnedResultBuffer() override { PinnedCache::Get().Release(ptr_, capacity_); }

 private:
  void* ptr_;
  size_t capacity_;
};

constexpr int64_t kPinnedResultMinBytes = 1 << 20;

// the context's non-blocking stream (or a caller's cudaStream_t*) behind arrow::Device::Stream
class B200Stream : public arrow::Device::Stream {
 public:
  B200Stream(B2Context* ctx, void* stream_ptr, release_fn_t release) : arrow::Device::Stream(stream_ptr, std::move(release)), ctx_(ctx) {}
  void* cuda_stream() const { return stream_ ? *reinterpret_cast<void* const*>(stream_.get()) : nullptr; }
  Status WaitEvent(const arrow::Device::SyncEvent& ev) override {
    void* raw = const_cast<arrow::Device::SyncEvent&>(ev).get_raw();
    if (!raw) return Status::OK();
    B200_RETURN_NOT_OK(b2_stream_wait_event(ctx_, cuda_stream(), *reinterpret_cast<void**>(raw)));
    return Status::OK();
  }
  Status Synchronize() const override {
    B200_RETURN_NOT_OK(b2_sync(ctx_, cuda_stream()));
    return Status::OK();
  }

 private:
  B2Context* ctx_;
};

// cudaEvent_t* behind arrow::Device::SyncEvent (the layout ArrowDeviceArray.sync_event has for CUDA)
class B200SyncEvent : public arrow::Device::SyncEvent {
 public:
  B200SyncEvent(B2Context* ctx, void* event_ptr, release_fn_t release) : arrow::Device::SyncEvent(event_ptr, std::move(release)), ctx_(ctx) {}
  Status Wait() override {
    B200_RETURN_NOT_OK(b2_event_synchronize(*reinterpret_cast<void**>(get_raw())));
    return Status::OK();
  }
  Status Record(const arrow::Device::Stream& s) override {
    const void* raw = s.get_raw();
    void* stream = raw ? *reinterpret_cast<void* const*>(raw) : nullptr;
    B200_RETURN_NOT_OK(b2_event_record(ctx_, *reinterpret_cast<void**>(get_raw()), stream));
    return Status::OK();
  }

 private:
  B2Context* ctx_;
};

}  // namespace

Result<std::shared_ptr<arrow::Device::Stream>> B200Device::MakeStream() {
  void** holder = new void*(b2_context_stream(ctx_));
  return std::shared_ptr<arrow::Device::Stream>(new B200Stream(ctx_, holder, [](void* p) { delete reinterpret_cast<void**>(p); }));
}

Result<std::shared_ptr<arrow::Device::Stream>> B200Device::WrapStream(void* cuda_stream_ptr, arrow::Device::Stream::release_fn_t release) {
  return std::shared_ptr<arrow::Device::Stream>(new B200Stream(ctx_, cuda_stream_ptr, release ? std::move(release) : [](void*) {}));
}

Result<std::shared_ptr<arrow::Device::SyncEvent>> B200MemoryManager::MakeDeviceSyncEvent() {
  void* ev = nullptr;
  B200_RETURN_NOT_OK(b2_event_create(context(), &ev));
  void** holder = new void*(ev);
  return std::shared_ptr<arrow::Device::SyncEvent>(new B200SyncEvent(context(), holder, [](void* p) {
    void** h = reinterpret_cast<void**>(p);
    b2_event_destroy(*h);
    delete h;
  }));
}

Result<std::shared_ptr<arrow::Device::SyncEvent>> B200MemoryManager::WrapDeviceSyncEvent(
    void* sync_event, arrow::Device::SyncEvent::release_fn_t release_sync_event) {
  return std::shared_ptr<arrow::Device::SyncEvent>(
      new B200SyncEvent(context(), sync_event, release_sync_event ? std::move(release_sync_event) : [](void*) {}));
}

Status ExportDeviceArray(const arrow::Array& array, const std::shared_ptr<arrow::MemoryManager>& mm, struct ArrowDeviceArray* out,
                         struct ArrowSchema* out_schema) {
  ARROW_ASSIGN_OR_RAISE(auto sync, mm->MakeDeviceSyncEvent());
  ARROW_ASSIGN_OR_RAISE(auto stream, mm->device()->MakeStream());
  ARROW_RETURN_NOT_OK(sync->Record(*stream));
  return arrow::ExportDeviceArray(array, std::move(sync), out, out_schema);
}

Result<std::shared_ptr<arrow::Array>> ImportDeviceArray(struct ArrowDeviceArray* array, std::shared_ptr<arrow::DataType> type,
                                                        const std::shared_ptr<arrow::MemoryManager>& mm) {
  if (array->device_type != ARROW_DEVICE_CUDA && array->device_type != ARROW_DEVICE_CUDA_MANAGED)
    return Status::Invalid("arrow_b200::ImportDeviceArray: device_type ", array->device_type, " is not CUDA memory");
  void* event_ptr = array->sync_event;
  void* event = event_ptr ? *reinterpret_cast<void**>(event_ptr) : nullptr;
  auto* b2mm = static_cast<B200MemoryManager*>(mm.get());
  // order our stream behind the producer BEFORE the structure (and with it the event) can be released
  if (event) B200_RETURN_NOT_OK(b2_stream_wait_event(b2mm->context(), nullptr, event));
  auto mapper = [mm](ArrowDeviceType, int64_t) -> Result<std::shared_ptr<arrow::MemoryManager>> { return mm; };
  ARROW_ASSIGN_OR_RAISE(auto imported, arrow::ImportDeviceArray(array, std::move(type), mapper));
  // a device array must carry an explicit null_count (ArrayData::GetNullCount would read the bitmap on the host)
  if (imported->data()->null_count.load() == arrow::kUnknownNullCount && imported->data()->buffers[0] == nullptr)
    imported->data()->null_count = 0;
  return imported;
}

namespace {
// orders the context stream behind the sync events of every buffer of every batch it hands out
class OrderedDeviceReader : public arrow::RecordBatchReader {
 public:
  OrderedDeviceReader(std::shared_ptr<arrow::RecordBatchReader> inner, std::shared_ptr<arrow::MemoryManager> mm)
      : inner_(std::move(inner)), mm_(std::move(mm)) {}
  std::shared_ptr<arrow::Schema> schema() const override { return inner_->schema(); }
  arrow::DeviceAllocationType device_type() const override { return arrow::DeviceAllocationType::kCUDA; }
  Status ReadNext(std::shared_ptr<arrow::RecordBatch>* out) override {
    ARROW_RETURN_NOT_OK(inner_->ReadNext(out));
    if (!*out) return Status::OK();
    ARROW_ASSIGN_OR_RAISE(auto stream, mm_->device()->MakeStream());
    for (const auto& col : (*out)->column_data()) {
      for (const auto& buf : col->buffers) {
        if (!buf) continue;
        if (auto ev = buf->device_sync_event()) ARROW_RETURN_NOT_OK(stream->WaitEvent(*ev));
      }
      // a device array must carry an explicit null_count (GetNullCount would read the bitmap on the host)
      if (col->null_count.load() == arrow::kUnknownNullCount && col->buffers[0] == nullptr) col->null_count = 0;
    }
    return Status::OK();
  }
  Status Close() override { return inner_->Close(); }

 private:
  std::shared_ptr<arrow::RecordBatchReader> inner_;
  std::shared_ptr<arrow::MemoryManager> mm_;
};
}  // namespace

Result<std::shared_ptr<arrow::RecordBatchReader>> ImportDeviceRecordBatchReader(struct ArrowDeviceArrayStream* stream,
                                                                                const std::shared_ptr<arrow::MemoryManager>& mm) {
  if (stream->device_type != ARROW_DEVICE_CUDA && stream->device_type != ARROW_DEVICE_CUDA_MANAGED)
    return Status::Invalid("arrow_b200::ImportDeviceRecordBatchReader: device_type ", stream->device_type, " is not CUDA memory");
  auto mapper = [mm](ArrowDeviceType, int64_t) -> Result<std::shared_ptr<arrow::MemoryManager>> { return mm; };
  ARROW_ASSIGN_OR_RAISE(auto inner, arrow::ImportDeviceRecordBatchReader(stream, mapper));
  return std::shared_ptr<arrow::RecordBatchReader>(new OrderedDeviceReader(std::move(inner), mm));
}

Result<std::shared_ptr<B200Device>> B200Device::Make(int device_number) {
  B2Context* ctx = nullptr;
  B200_RETURN_NOT_OK(b2_context_create(device_number, &ctx));
  return std::shared_ptr<B200Device>(new B200Device(device_number, ctx));
}

B200Device::~B200Device() { b2_context_destroy(ctx_); }

std::string B200Device::ToString() const { return "B200Device(device " + std::to_string(device_number_) + ")"; }

bool B200Device::Equals(const arrow::Device& other) const {
  if (other.type_name() != type_name()) return false;
  return static_cast<const B200Device&>(other).device_number_ == device_number_;
}

std::shared_ptr<arrow::MemoryManager> B200Device::default_memory_manager() {
  auto mm = mm_.lock();
  if (!mm) {
    mm = std::make_shared<B200MemoryManager>(shared_from_this());
    mm_ = mm;
  }
  return mm;
}

Result<std::shared_ptr<arrow::io::RandomAccessFile>> B200MemoryManager::GetBufferReader(std::shared_ptr<arrow::Buffer>) {
  return Status::NotImplemented("B200MemoryManager::GetBufferReader");
}
Result<std::shared_ptr<arrow::io::OutputStream>> B200MemoryManager::GetBufferWriter(std::shared_ptr<arrow::Buffer>) {
  return Status::NotImplemented("B200MemoryManager::GetBufferWriter");
}

Result<std::unique_ptr<arrow::Buffer>> B200MemoryManager::AllocateBuffer(int64_t size) {
  void* p = nullptr;
  B200_RETURN_NOT_OK(b2_alloc(context(), static_cast<size_t>(size > 0 ? size : 1), &p));
  return std::unique_ptr<arrow::Buffer>(new B200Buffer(static_cast<uint8_t*>(p), size, shared_from_this(), context()));
}

std::shared_ptr<arrow::Buffer> B200MemoryManager::Adopt(const void* ptr, int64_t size) {
  if (!ptr) return nullptr;
  return std::make_shared<B200Buffer>(static_cast<uint8_t*>(const_cast<void*>(ptr)), size, shared_from_this(), context());
}

Result<std::shared_ptr<arrow::Buffer>> B200MemoryManager::CopyBufferFrom(const std::shared_ptr<arrow::Buffer>& buf,
                                                                        const std::shared_ptr<arrow::MemoryManager>& from) {
  ARROW_ASSIGN_OR_RAISE(auto out, CopyNonOwnedFrom(*buf, from));
  return std::shared_ptr<arrow::Buffer>(std::move(out));
}

Result<std::unique_ptr<arrow::Buffer>> B200MemoryManager::CopyNonOwnedFrom(const arrow::Buffer& buf,
                                                                          const std::shared_ptr<arrow::MemoryManager>& from) {
  if (!from->is_cpu()) return nullptr;  // unsupported source: let arrow try other routes
  ARROW_ASSIGN_OR_RAISE(auto out, AllocateBuffer(buf.size()));
  B200_RETURN_NOT_OK(b2_memcpy_h2d(context(), reinterpret_cast<void*>(out->address()), buf.data(),
                                   static_cast<size_t>(buf.size()), nullptr));
  B200_RETURN_NOT_OK(b2_sync(context(), nullptr));
  return out;
}

Result<std::shared_ptr<arrow::Buffer>> B200MemoryManager::CopyBufferTo(const std::shared_ptr<arrow::Buffer>& buf,
                                                                      const std::shared_ptr<arrow::MemoryManager>& to) {
  ARROW_ASSIGN_OR_RAISE(auto out, CopyNonOwnedTo(*buf, to));
  return std::shared_ptr<arrow::Buffer>(std::move(out));
}

Result<std::unique_ptr<arrow::Buffer>> B200MemoryManager::CopyNonOwnedTo(const arrow::Buffer& buf,
                                                                        const std::shared_ptr<arrow::MemoryManager>& to) {
  if (!to->is_cpu()) return nullptr;
  std::unique_ptr<arrow::Buffer> out;
  if (buf.size() >= kPinnedResultMinBytes && to == arrow::default_cpu_memory_manager()) {
    size_t capacity = 0;
    ARROW_ASSIGN_OR_RAISE(void* p, PinnedCache::Get().Acquire(static_cast<size_t>(buf.size()), &capacity));
    out.reset(new PinnedResultBuffer(p, buf.size(), capacity));
  } else {
    ARROW_ASSIGN_OR_RAISE(out, to->AllocateBuffer(buf.size()));
  }
  B200_RETURN_NOT_OK(b2_memcpy_d2h(context(), out->mutable_data(), reinterpret_cast<const void*>(buf.address()),
                                   static_cast<size_t>(buf.size()), nullptr));
  B200_RETURN_NOT_OK(b2_sync(context(), nullptr));
  return out;
}

bool IsOnDevice(const arrow::ArrayData& data) {
  for (const auto& b : data.buffers)
    if (b && !b->is_cpu()) return true;
  for (const auto& c : data.child_data)
    if (c && IsOnDevice(*c)) return true;
  if (data.dictionary && IsOnDevice(*data.dictionary)) return true;
  return false;
}

namespace {
Result<std::shared_ptr<arrow::ArrayData>> CopyData(const arrow::ArrayData& in, const std::shared_ptr<arrow::MemoryManager>& to) {
  auto out = std::make_shared<arrow::ArrayData>(in.type, in.length, in.null_count.load(), in.offset);
  out->buffers.resize(in.buffers.size());
  for (size_t i = 0; i < in.buffers.size(); ++i) {
    if (in.buffers[i]) {
      ARROW_ASSIGN_OR_RAISE(out->buffers[i], arrow::MemoryManager::CopyBuffer(in.buffers[i], to));
    }
  }
  for (const auto& child : in.child_data) {
    ARROW_ASSIGN_OR_RAISE(auto c, CopyData(*child, to));
    out->child_data.push_back(std::move(c));
  }
  if (in.dictionary) {
    ARROW_ASSIGN_OR_RAISE(out->dictionary, CopyData(*in.dictionary, to));
  }
  return out;
}
}  // namespace

Result<std::shared_ptr<arrow::ArrayData>> ToDevice(const arrow::ArrayData& host, const std::shared_ptr<arrow::MemoryManager>& mm) {
  // a device array must carry an explicit null_count: ArrayData::GetNullCount would
  // dereference a null bitmap pointer (array/data.cc:214-226)
  ARROW_ASSIGN_OR_RAISE(auto out, CopyData(host, mm));
  out->null_count = host.GetNullCount();
  return out;
}

Result<std::shared_ptr<arrow::ArrayData>> ToHost(const arrow::ArrayData& device) {
  return CopyData(device, arrow::default_cpu_memory_manager());
}

}  // namespace arrow_b200
