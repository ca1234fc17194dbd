// scalar_boolean.cu -- boolean logic and validity predicates over bit-packed columns.
//
// Replaces (the kernels an Expression filter is made of):
//   and / or / xor / and_not / invert            kernels/scalar_boolean.cc:30-270 (AndOp, OrOp, XorOp, AndNotOp, InvertOp)
//   and_kleene / or_kleene / and_not_kleene      kernels/scalar_boolean.cc:110-230 (KleeneAndOp, KleeneOrOp, KleeneAndNotOp)
//   is_valid / is_null / true_unless_null / is_nan   kernels/scalar_validity.cc:35-260 (NullOptions{nan_is_null})
// Semantics kept: the plain ops are null when either operand is null (validity = AND of the input
// validities, data = op on the data bits of EVERY slot); the Kleene ops know that `false AND x` is
// false and `true OR x` is true whatever x is; is_valid / is_null / is_nan never return nulls.
//
// B200 design: one thread produces one aligned 64-bit output word (data and, when needed, validity)
// from funnel-shifted input words, so every operand may start at any bit offset; a scalar operand is a
// broadcast word.  1/8 B per row per bitmap touched: pure HBM streaming.
#include "bitmap.h"

namespace b2 {

struct BoolOperand {
  BitmapReader data, valid;
  uint64_t scalar_data, scalar_valid;  // used when is_scalar
  bool is_scalar, has_valid;
  __device__ __forceinline__ uint64_t d(int64_t w) const { return is_scalar ? scalar_data : data.word(w); }
  __device__ __forceinline__ uint64_t v(int64_t w) const {
    return is_scalar ? scalar_valid : (has_valid ? valid.word(w) : ~0ull);
  }
};

__global__ void __launch_bounds__(kBlock) boolean_kernel(int op, BoolOperand a, BoolOperand b, int64_t n, int64_t nwords,
                                                         uint64_t* out_data, uint64_t* out_valid, int64_t* valid_count) {
  int64_t local = 0;
  for (int64_t w = blockIdx.x * (int64_t)kBlock + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * kBlock) {
    const int64_t rem = n - (w << 6);
    const uint64_t keep = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
    const uint64_t ad = a.d(w), av = a.v(w);
    uint64_t bd = 0, bv = ~0ull;
    if (op != B2_BOOL_INVERT) {
      bd = b.d(w);
      bv = b.v(w);
    }
    uint64_t d, v;
    switch (op) {
      case B2_BOOL_AND: d = ad & bd; v = av & bv; break;
      case B2_BOOL_OR: d = ad | bd; v = av & bv; break;
      case B2_BOOL_XOR: d = ad ^ bd; v = av & bv; break;
      case B2_BOOL_AND_NOT: d = ad & ~bd; v = av & bv; break;
      case B2_BOOL_AND_KLEENE: d = ad & bd; v = (av & bv) | (av & ~ad) | (bv & ~bd); break;
      case B2_BOOL_OR_KLEENE: d = ad | bd; v = (av & bv) | (av & ad) | (bv & bd); break;
      case B2_BOOL_AND_NOT_KLEENE: d = ad & ~bd; v = (av & bv) | (av & ~ad) | (bv & bd); break;
      default: d = ~ad; v = av; break;  // invert
    }
    d &= keep;
    v &= keep;
    out_data[w] = d;
    if (out_valid) out_valid[w] = v;
    local += __popcll(v);
  }
  if (valid_count) {
    int64_t s = block_sum<kBlock>(local);
    if (threadIdx.x == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(valid_count), (unsigned long long)s);
  }
}

// is_valid / is_null / true_unless_null on the validity bitmap alone
__global__ void __launch_bounds__(kBlock) validity_kernel(int op, BitmapReader valid, int64_t n, int64_t nwords,
                                                          uint64_t* out_data) {
  for (int64_t w = blockIdx.x * (int64_t)kBlock + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * kBlock) {
    const int64_t rem = n - (w << 6);
    const uint64_t keep = rem >= 64 ? ~0ull : ((1ull << rem) - 1ull);
    const uint64_t v = valid.word(w);
    uint64_t d = op == B2_IS_NULL ? ~v : (op == B2_IS_VALID ? v : ~0ull);
    out_data[w] = d & keep;
  }
}

// is_nan, and is_null with nan_is_null: one value per lane, packed with a ballot (coalesced value reads)
template <typename T>
__global__ void __launch_bounds__(kBlock) nan_kernel(int op, const T* __restrict__ values, BitmapReader valid, int64_t n,
                                                     uint32_t* out_data) {
  const int64_t nw = (n + 31) >> 5;
  for (int64_t w = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5; w < nw; w += ((int64_t)gridDim.x * kBlock) >> 5) {
    const int64_t i = (w << 5) + lane_id();
    bool r = false;
    if (i < n) {
      const bool ok = valid.bit(i);
      const T x = __ldcs(values + i);
      const bool is_nan = x != x;
      r = op == B2_IS_NAN ? is_nan : (!ok || is_nan);  // is_null(nan_is_null): values under nulls are ignored
    }
    const unsigned word = __ballot_sync(0xffffffffu, r);
    if (lane_id() == 0) out_data[w] = word;
  }
}

static int make_bool_operand(const B2Value* v, int64_t* length, BoolOperand* o, const char* who) {
  memset(o, 0, sizeof(*o));
  if (v->array) {
    const B2Array* a = v->array;
    if (a->type != B2_BOOL) return set_error(B2_TYPE_ERROR, "%s: operands must be boolean (type id %d)", who, a->type);
    if (a->length < 0 || a->offset < 0) return set_error(B2_INVALID, "negative length/offset");
    if (*length >= 0 && *length != a->length) return set_error(B2_INVALID, "%s: array operands differ in length", who);
    *length = a->length;
    o->data = BitmapReader(a->data, a->offset, a->length);
    o->has_valid = a->validity && a->null_count != 0;
    o->valid = BitmapReader(o->has_valid ? a->validity : nullptr, a->offset, a->length);
    o->is_scalar = false;
  } else if (v->scalar) {
    if (v->scalar->type != B2_BOOL) return set_error(B2_TYPE_ERROR, "%s: operands must be boolean", who);
    o->is_scalar = true;
    o->scalar_valid = v->scalar->is_valid ? ~0ull : 0ull;
    o->scalar_data = (v->scalar->is_valid && (v->scalar->bits & 1)) ? ~0ull : 0ull;
    o->has_valid = !v->scalar->is_valid;
  } else {
    return set_error(B2_INVALID, "%s: empty operand", who);
  }
  return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_boolean(B2Context* ctx, int op, const B2Value* left, const B2Value* right, B2Array* out, void* stream) {
  if (!ctx || !left || !out) return set_error(B2_INVALID, "b2_boolean: null argument");
  if (op < B2_BOOL_AND || op > B2_BOOL_INVERT) return set_error(B2_INVALID, "b2_boolean: unknown op %d", op);
  if (op != B2_BOOL_INVERT && !right) return set_error(B2_INVALID, "b2_boolean: binary op needs two operands");
  cudaStream_t s = ctx->pick(stream);
  B2_CUDA(cudaSetDevice(ctx->device));
  int64_t n = -1;
  BoolOperand a, b;
  B2_RETURN_NOT_OK(make_bool_operand(left, &n, &a, "b2_boolean"));
  if (op != B2_BOOL_INVERT) B2_RETURN_NOT_OK(make_bool_operand(right, &n, &b, "b2_boolean"));
  else memset(&b, 0, sizeof(b));
  if (n < 0) return set_error(B2_NOT_IMPLEMENTED, "b2_boolean: at least one operand must be an array");
  Temp data(ctx, s), bits(ctx, s);
  B2_RETURN_NOT_OK(data.alloc(bitmap_alloc_bytes(n)));
  if (n == 0) {
    fill_out(out, B2_BOOL, 0, 0, nullptr, data.release());
    return B2_OK;
  }
  const bool may_null = a.has_valid || (op != B2_BOOL_INVERT && b.has_valid);
  const int64_t nwords = bitmap_words64(n);
  B2_CUDA(cudaMemsetAsync(static_cast<char*>(data.ptr) + nwords * 8, 0, 8, s));
  int64_t nulls = 0;
  if (may_null) {
    B2_RETURN_NOT_OK(bits.alloc(bitmap_alloc_bytes(n)));
    B2_CUDA(cudaMemsetAsync(static_cast<char*>(bits.ptr) + nwords * 8, 0, 8, s));
    ScalarSlot slot(ctx);
    B2_RETURN_NOT_OK(slot.zero(s));
    boolean_kernel<<<grid_for(nwords, kBlock, kSMs * 8), kBlock, 0, s>>>(op, a, b, n, nwords, data.as<uint64_t>(),
                                                                        bits.as<uint64_t>(), slot.dev());
    B2_LAUNCHED();
    B2_RETURN_NOT_OK(slot.fetch(s));
    nulls = n - slot.host()[0];
  } else {
    boolean_kernel<<<grid_for(nwords, kBlock, kSMs * 8), kBlock, 0, s>>>(op, a, b, n, nwords, data.as<uint64_t>(), nullptr,
                                                                        nullptr);
    B2_LAUNCHED();
  }
  fill_out(out, B2_BOOL, n, nulls, nulls ? bits.release() : nullptr, data.release());
  return B2_OK;
}

extern "C" int b2_validity(B2Context* ctx, int op, const B2Array* in, int nan_is_null, B2Array* out, void* stream) {
  if (!ctx || !in || !out) return set_error(B2_INVALID, "b2_validity: null argument");
  if (op < B2_IS_VALID || op > B2_IS_NAN) return set_error(B2_INVALID, "b2_validity: unknown op %d", op);
  if (in->length < 0 || in->offset < 0) return set_error(B2_INVALID, "negative length/offset");
  const bool is_float = in->type == B2_FLOAT || in->type == B2_DOUBLE;
  if (op == B2_IS_NAN && !is_float)
    return set_error(B2_NOT_IMPLEMENTED, "is_nan: only float32 / float64 arrays (type id %d)", in->type);
  cudaStream_t s = ctx->pick(stream);
  B2_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = in->length;
  Temp data(ctx, s), bits(ctx, s);
  B2_RETURN_NOT_OK(data.alloc(bitmap_alloc_bytes(n)));
  if (n == 0) {
    fill_out(out, B2_BOOL, 0, 0, nullptr, data.release());
    return B2_OK;
  }
  const int64_t nwords = bitmap_words64(n);
  B2_CUDA(cudaMemsetAsync(static_cast<char*>(data.ptr) + (nwords - 1) * 8, 0, 16, s));
  const bool has_valid = in->validity && in->null_count != 0;
  BitmapReader valid(has_valid ? in->validity : nullptr, in->offset, n);
  if (op == B2_IS_NAN || (op == B2_IS_NULL && nan_is_null && is_float)) {
    const int grid = grid_for(n, kBlock * 4, kSMs * 16);
    if (in->type == B2_FLOAT)
      nan_kernel<float><<<grid, kBlock, 0, s>>>(op, static_cast<const float*>(in->data) + in->offset, valid, n, data.as<uint32_t>());
    else
      nan_kernel<double><<<grid, kBlock, 0, s>>>(op, static_cast<const double*>(in->data) + in->offset, valid, n, data.as<uint32_t>());
    B2_LAUNCHED();
  } else {
    validity_kernel<<<grid_for(nwords, kBlock, kSMs * 8), kBlock, 0, s>>>(op, valid, n, nwords, data.as<uint64_t>());
    B2_LAUNCHED();
  }
  // is_nan of a null slot is null (ScalarUnary with INTERSECTION); true_unless_null keeps the input validity
  void* validity = nullptr;
  int64_t nulls = 0;
  if ((op == B2_TRUE_UNLESS_NULL || op == B2_IS_NAN) && has_valid)
    B2_RETURN_NOT_OK(make_validity(ctx, in, nullptr, n, &validity, &nulls, s));
  fill_out(out, B2_BOOL, n, nulls, validity, data.release());
  return B2_OK;
}
