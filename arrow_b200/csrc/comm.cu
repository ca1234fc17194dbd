// comm.cu -- the exchange step of the multi-GPU hash-aggregate and SortIndices behind the C-ABI
// (b2_comm_{init,all_to_all_v,destroy}; section 8e: "one all-to-all-v via
// grouped ncclSend/ncclRecv preceded by an all-gather of the P x P counts").
//
// The reference is single-process; its per-thread analogue of this exchange is GroupByNode::Merge
// (acero/groupby_aggregate_node.cc:255-298).  One process per GPU: the host application creates the
// NCCL unique id on rank 0 (b2_comm_unique_id), distributes the 128 bytes by whatever means it has
// (MPI, a torch.distributed store, a file), and every rank calls b2_comm_init.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that libarrow_b200.so has no link-time
// dependency on it: a process that already loaded NCCL (PyTorch does) shares that copy; single-GPU
// users never need the library.  All transfers are stream-ordered on the caller's stream; over
// NVLink 5 / NVSwitch every peer pair has full bandwidth, so the P-1 sends and receives of one
// all-to-all-v are issued as ONE group and proceed concurrently.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "context.h"

namespace b2 {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.handle) return B2_OK;
  void* h = nullptr;
  const char* env = getenv("B2_NCCL_LIBRARY");
  if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  // RTLD_NOLOAD first: reuse the copy the process already mapped (e.g. the one PyTorch bundles)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_error(B2_IO_ERROR, "b2_comm: cannot load NCCL (libnccl.so.2): %s", dlerror());
  NcclApi api;
  api.handle = h;
#define B2_NCCL_SYM(field, name)                                                              \
  *reinterpret_cast<void**>(&api.field) = dlsym(h, name);                                     \
  if (!api.field) return set_error(B2_IO_ERROR, "b2_comm: NCCL symbol %s not found", name)
  B2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  B2_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  B2_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  B2_NCCL_SYM(GroupStart, "ncclGroupStart");
  B2_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  B2_NCCL_SYM(Send, "ncclSend");
  B2_NCCL_SYM(Recv, "ncclRecv");
  B2_NCCL_SYM(AllGather, "ncclAllGather");
  B2_NCCL_SYM(AllReduce, "ncclAllReduce");
  B2_NCCL_SYM(GetErrorString, "ncclGetErrorString");
  B2_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef B2_NCCL_SYM
  g_nccl = api;
  return B2_OK;
}

#define B2_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    ncclResult_t _r = (expr);                                                                           \
    if (_r != ncclSuccess)                                                                              \
      return ::b2::set_error(B2_IO_ERROR, "NCCL error at %s:%d: %s", __FILE__, __LINE__,                \
                             ::b2::g_nccl.GetErrorString(_r));                                          \
  } while (0)

}  // namespace b2

using namespace b2;

struct B2Comm {
  B2Context* ctx;
  ncclComm_t comm;
  int rank, world;
};

extern "C" {

int b2_comm_unique_id(uint8_t* out_id) {
  if (!out_id) return set_error(B2_INVALID, "b2_comm_unique_id: null argument");
  B2_RETURN_NOT_OK(load_nccl());
  static_assert(sizeof(ncclUniqueId) == B2_COMM_ID_BYTES, "B2_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  ncclUniqueId id;
  B2_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id, &id, sizeof(id));
  return B2_OK;
}

int b2_comm_init(B2Context* ctx, int rank, int world, const uint8_t* id, B2Comm** out) {
  if (!ctx || !id || !out) return set_error(B2_INVALID, "b2_comm_init: null argument");
  if (world < 1 || rank < 0 || rank >= world) return set_error(B2_INVALID, "b2_comm_init: bad rank %d of %d", rank, world);
  B2_RETURN_NOT_OK(load_nccl());
  B2_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm;
  B2_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  B2Comm* c = new B2Comm{ctx, comm, rank, world};
  *out = c;
  return B2_OK;
}

void b2_comm_destroy(B2Comm* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

int b2_comm_rank(const B2Comm* c) { return c ? c->rank : -1; }
int b2_comm_world(const B2Comm* c) { return c ? c->world : 0; }

int b2_comm_nccl_version(int* out) {
  if (!out) return set_error(B2_INVALID, "b2_comm_nccl_version: null argument");
  B2_RETURN_NOT_OK(load_nccl());
  B2_NCCL(g_nccl.GetVersion(out));
  return B2_OK;
}

// Several all_to_all_v calls (one per column of a partitioned batch) bracketed by group_start / group_end
// become ONE NCCL launch: the columns travel together without being packed into a staging buffer.
int b2_comm_group_start(B2Comm* c) {
  if (!c) return set_error(B2_INVALID, "b2_comm_group_start: null argument");
  B2_CUDA(cudaSetDevice(c->ctx->device));
  B2_NCCL(g_nccl.GroupStart());
  return B2_OK;
}

int b2_comm_group_end(B2Comm* c) {
  if (!c) return set_error(B2_INVALID, "b2_comm_group_end: null argument");
  B2_CUDA(cudaSetDevice(c->ctx->device));
  B2_NCCL(g_nccl.GroupEnd());
  return B2_OK;
}

int b2_comm_all_gather(B2Comm* c, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
  if (!c || !send || !recv) return set_error(B2_INVALID, "b2_comm_all_gather: null argument");
  if (bytes_per_rank < 0) return set_error(B2_INVALID, "negative size");
  cudaStream_t s = c->ctx->pick(stream);
  B2_CUDA(cudaSetDevice(c->ctx->device));
  B2_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, c->comm, s));
  return B2_OK;
}

int b2_comm_all_reduce_i64(B2Comm* c, const void* send, void* recv, int64_t count, int op, void* stream) {
  if (!c || !send || !recv) return set_error(B2_INVALID, "b2_comm_all_reduce_i64: null argument");
  if (op != B2_COMM_SUM && op != B2_COMM_MAX && op != B2_COMM_MIN) return set_error(B2_INVALID, "bad reduction op %d", op);
  cudaStream_t s = c->ctx->pick(stream);
  B2_CUDA(cudaSetDevice(c->ctx->device));
  const ncclRedOp_t rop = op == B2_COMM_SUM ? ncclSum : (op == B2_COMM_MAX ? ncclMax : ncclMin);
  B2_NCCL(g_nccl.AllReduce(send, recv, (size_t)count, ncclInt64, rop, c->comm, s));
  return B2_OK;
}

int b2_comm_all_to_all_v(B2Comm* c, const void* send, const int64_t* send_offsets, const int64_t* send_bytes, void* recv,
                         const int64_t* recv_offsets, const int64_t* recv_bytes, void* stream) {
  if (!c || !send_offsets || !send_bytes || !recv_offsets || !recv_bytes)
    return set_error(B2_INVALID, "b2_comm_all_to_all_v: null argument");
  cudaStream_t s = c->ctx->pick(stream);
  B2_CUDA(cudaSetDevice(c->ctx->device));
  for (int p = 0; p < c->world; ++p)
    if (send_bytes[p] < 0 || recv_bytes[p] < 0 || send_offsets[p] < 0 || recv_offsets[p] < 0)
      return set_error(B2_INVALID, "b2_comm_all_to_all_v: negative size/offset for peer %d", p);
  // this rank's own chunk never leaves the device: a plain D2D copy
  if (send_bytes[c->rank] != recv_bytes[c->rank])
    return set_error(B2_INVALID, "b2_comm_all_to_all_v: self chunk sizes differ (%lld vs %lld)", (long long)send_bytes[c->rank],
                     (long long)recv_bytes[c->rank]);
  if (send_bytes[c->rank] > 0)
    B2_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + recv_offsets[c->rank], static_cast<const char*>(send) + send_offsets[c->rank],
                            (size_t)send_bytes[c->rank], cudaMemcpyDeviceToDevice, s));
  if (c->world == 1) return B2_OK;
  B2_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    if (send_bytes[p] > 0)
      B2_NCCL(g_nccl.Send(static_cast<const char*>(send) + send_offsets[p], (size_t)send_bytes[p], ncclUint8, p, c->comm, s));
    if (recv_bytes[p] > 0)
      B2_NCCL(g_nccl.Recv(static_cast<char*>(recv) + recv_offsets[p], (size_t)recv_bytes[p], ncclUint8, p, c->comm, s));
  }
  B2_NCCL(g_nccl.GroupEnd());
  return B2_OK;
}

}  // extern "C"
