// scalar_arith.cu -- add / subtract / multiply / divide (+ _checked variants).
//
// Replaces the reference's arithmetic applicators and op functors:
//   ScalarBinary<..>::{ArrayArray,ArrayScalar,ScalarArray}   kernels/codegen_internal.h:813-874
//   ScalarBinaryNotNull* (checked ops, divide: null slots are never checked)  :878-976
//   Add/AddChecked/Subtract/SubtractChecked/Multiply/MultiplyChecked/Divide/DivideChecked
//                                                  kernels/base_arithmetic_internal.h:44-423
// Semantics kept: unchecked signed-integer ops wrap (computed in unsigned arithmetic),
// floats are a single IEEE operation (bit-exact; no FMA contraction possible, no
// flush-to-zero), checked ops raise Invalid("overflow") / Invalid("divide by zero") for
// the first offending VALID slot, integer INT_MIN / -1 is 0 for `divide`, overflow for
// `divide_checked`.  Both operands arrive with the dispatched common type
// (ArithmeticFunction::DispatchBest, kernels/scalar_arithmetic.cc:734-781).
//
// Roofline: HBM-bound; algorithmic bytes/row = 3*sizeof(T) + 3/8 with two nullable
// inputs (Add f32: 12.375 B/row).
#include <limits>
#include <type_traits>

#include "bitmap.h"
#include "elementwise.cuh"

namespace b2 {

template <typename T>
using U = typename std::make_unsigned<T>::type;

template <typename T, int OP>
struct ArithOp {
  BitmapReader valid;  // output validity (offset 0) -- only consulted on a failure
  ErrorCell err;
  __device__ __forceinline__ T operator()(T l, T r, int64_t i) const {
    constexpr bool kInt = std::is_integral<T>::value;
    T res;
    bool bad = false;
    if constexpr (!kInt) {
      if constexpr (OP == B2_ADD || OP == B2_ADD_CHECKED) res = l + r;
      else if constexpr (OP == B2_SUBTRACT || OP == B2_SUBTRACT_CHECKED) res = l - r;
      else if constexpr (OP == B2_MULTIPLY || OP == B2_MULTIPLY_CHECKED) res = l * r;
      else if constexpr (OP == B2_DIVIDE) res = l / r;
      else {  // DivideChecked, base_arithmetic_internal.h:406-414
        bad = (r == 0);
        res = bad ? T(0) : l / r;
      }
    } else {
      using UT = U<T>;
      using W = typename std::conditional<(sizeof(T) < 4), uint32_t, UT>::type;
      if constexpr (OP == B2_ADD) {
        res = static_cast<T>(static_cast<W>(static_cast<UT>(l)) + static_cast<W>(static_cast<UT>(r)));
      } else if constexpr (OP == B2_SUBTRACT) {
        res = static_cast<T>(static_cast<W>(static_cast<UT>(l)) - static_cast<W>(static_cast<UT>(r)));
      } else if constexpr (OP == B2_MULTIPLY) {
        res = static_cast<T>(static_cast<W>(static_cast<UT>(l)) * static_cast<W>(static_cast<UT>(r)));
      } else if constexpr (OP == B2_ADD_CHECKED) {
        res = static_cast<T>(static_cast<W>(static_cast<UT>(l)) + static_cast<W>(static_cast<UT>(r)));
        if constexpr (std::is_signed<T>::value) bad = ((l ^ res) & (r ^ res)) < 0;
        else bad = res < l;
      } else if constexpr (OP == B2_SUBTRACT_CHECKED) {
        res = static_cast<T>(static_cast<W>(static_cast<UT>(l)) - static_cast<W>(static_cast<UT>(r)));
        if constexpr (std::is_signed<T>::value) bad = ((l ^ r) & (l ^ res)) < 0;
        else bad = l < r;
      } else if constexpr (OP == B2_MULTIPLY_CHECKED) {
        if constexpr (sizeof(T) < 8) {
          using Wide = typename std::conditional<std::is_signed<T>::value, int64_t, uint64_t>::type;
          Wide p = static_cast<Wide>(l) * static_cast<Wide>(r);
          res = static_cast<T>(p);
          bad = static_cast<Wide>(res) != p;
        } else if constexpr (std::is_signed<T>::value) {
          res = static_cast<T>(static_cast<UT>(l) * static_cast<UT>(r));
          long long hi = __mul64hi(static_cast<long long>(l), static_cast<long long>(r));
          bad = hi != (static_cast<long long>(res) >> 63);
        } else {
          res = l * r;
          bad = __umul64hi(l, r) != 0;
        }
      } else {  // DIVIDE / DIVIDE_CHECKED, base_arithmetic_internal.h:364-400
        bool overflow = false;
        if constexpr (std::is_signed<T>::value)
          overflow = (l == std::numeric_limits<T>::min() && r == T(-1));
        if (r == 0) {
          bad = true;
          res = 0;
        } else if (overflow) {
          bad = (OP == B2_DIVIDE_CHECKED);
          res = (OP == B2_DIVIDE_CHECKED) ? l : T(0);
        } else {
          res = l / r;
        }
      }
    }
    if (bad && valid.bit(i)) err.report(i);
    return res;
  }
};

template <typename T, int OP>
static int run_arith(const T* a, T sa, const T* b, T sb, T* out, int64_t n, const void* out_validity,
                     unsigned long long* d_err, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  bool vec_ok = aligned_to(out, 16) && (!a || aligned_to(a, 16)) && (!b || aligned_to(b, 16));
  ArithOp<T, OP> op;
  op.valid = BitmapReader(out_validity, 0, n);
  op.err.first_row = d_err;
  map2_kernel<T, T, V, ArithOp<T, OP>>
      <<<map_grid<T, T, V>(n), kBlock, 0, s>>>(a, sa, b, sb, out, n, vec_ok, op);
  B2_LAUNCHED();
  return B2_OK;
}

template <typename T>
static int run_arith_op(int op, const T* a, T sa, const T* b, T sb, T* out, int64_t n,
                        const void* v, unsigned long long* e, cudaStream_t s) {
  switch (op) {
    case B2_ADD: return run_arith<T, B2_ADD>(a, sa, b, sb, out, n, v, e, s);
    case B2_SUBTRACT: return run_arith<T, B2_SUBTRACT>(a, sa, b, sb, out, n, v, e, s);
    case B2_MULTIPLY: return run_arith<T, B2_MULTIPLY>(a, sa, b, sb, out, n, v, e, s);
    case B2_DIVIDE: return run_arith<T, B2_DIVIDE>(a, sa, b, sb, out, n, v, e, s);
    case B2_ADD_CHECKED: return run_arith<T, B2_ADD_CHECKED>(a, sa, b, sb, out, n, v, e, s);
    case B2_SUBTRACT_CHECKED: return run_arith<T, B2_SUBTRACT_CHECKED>(a, sa, b, sb, out, n, v, e, s);
    case B2_MULTIPLY_CHECKED: return run_arith<T, B2_MULTIPLY_CHECKED>(a, sa, b, sb, out, n, v, e, s);
    case B2_DIVIDE_CHECKED: return run_arith<T, B2_DIVIDE_CHECKED>(a, sa, b, sb, out, n, v, e, s);
    default: return set_error(B2_INVALID, "unknown arithmetic op %d", op);
  }
}

// operand unpacking shared with scalar_compare.cu
int unpack_operands(const B2Value* left, const B2Value* right, int* type, int64_t* length,
                    bool* null_scalar) {
  if (!left || !right) return set_error(B2_INVALID, "null operand");
  const B2Array* la = left->array;
  const B2Array* ra = right->array;
  if (!la && !left->scalar) return set_error(B2_INVALID, "left operand has neither array nor scalar");
  if (!ra && !right->scalar) return set_error(B2_INVALID, "right operand has neither array nor scalar");
  if (!la && !ra) return set_error(B2_INVALID, "at least one operand must be an array");
  int lt = la ? la->type : left->scalar->type;
  int rt = ra ? ra->type : right->scalar->type;
  if (lt != rt)
    return set_error(B2_TYPE_ERROR, "operand types differ (%d vs %d); dispatch must cast first", lt, rt);
  if (la && ra && la->length != ra->length)
    return set_error(B2_INVALID, "Array arguments must all be the same length");
  *type = lt;
  *length = la ? la->length : ra->length;
  *null_scalar = (!la && !left->scalar->is_valid) || (!ra && !right->scalar->is_valid);
  return B2_OK;
}

template <typename T>
static T scalar_bits(const B2Scalar* s) {
  T v{};
  if (s) memcpy(&v, &s->bits, sizeof(T));
  return v;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_binary_arith(B2Context* ctx, int op, const B2Value* left, const B2Value* right,
                               B2Array* out, void* stream) {
  if (!ctx || !out) return set_error(B2_INVALID, "b2_binary_arith: null argument");
  int type;
  int64_t n;
  bool null_scalar;
  B2_RETURN_NOT_OK(unpack_operands(left, right, &type, &n, &null_scalar));
  if (!type_is_numeric(type)) return set_error(B2_NOT_IMPLEMENTED, "arithmetic on type id %d", type);
  cudaStream_t s = ctx->pick(stream);
  B2_CUDA(cudaSetDevice(ctx->device));
  const B2Array* la = left->array;
  const B2Array* ra = right->array;
  const int w = type_width(type);

  Temp data(ctx, s);
  B2_RETURN_NOT_OK(data.alloc(static_cast<size_t>(n) * w));
  if (null_scalar) {
    // a null scalar operand nulls every slot (exec.cc:560-590 all-null propagation)
    Temp bits(ctx, s);
    B2_RETURN_NOT_OK(bits.alloc(bitmap_alloc_bytes(n)));
    B2_CUDA(cudaMemsetAsync(bits.ptr, 0, bitmap_alloc_bytes(n), s));
    B2_CUDA(cudaMemsetAsync(data.ptr, 0, static_cast<size_t>(n) * w + (n == 0), s));
    fill_out(out, type, n, n, n ? bits.release() : nullptr, data.release());
    return B2_OK;
  }
  void* validity = nullptr;
  int64_t null_count = 0;
  B2_RETURN_NOT_OK(make_validity(ctx, la, ra, n, &validity, &null_count, s));
  Temp vguard(ctx, s);
  vguard.ptr = validity;
  if (n > 0) {
    const bool can_fail = (op >= B2_ADD_CHECKED) || (op == B2_DIVIDE);
    ScalarSlot slot(ctx);
    unsigned long long* d_err = nullptr;
    if (can_fail) {
      if (!slot.ok()) return set_error(B2_UNKNOWN_ERROR, "no free scalar slot");
      B2_CUDA(cudaMemsetAsync(slot.dev(), 0xff, 8, s));
      d_err = reinterpret_cast<unsigned long long*>(slot.dev());
    }
    int st = B2_OK;
#define B2_ARITH_CASE(ID, T)                                                                   \
  case ID: {                                                                                   \
    const T* a = la ? static_cast<const T*>(la->data) + la->offset : nullptr;                  \
    const T* b = ra ? static_cast<const T*>(ra->data) + ra->offset : nullptr;                  \
    st = run_arith_op<T>(op, a, scalar_bits<T>(left->scalar), b, scalar_bits<T>(right->scalar), \
                         static_cast<T*>(data.ptr), n, validity, d_err, s);                    \
    break;                                                                                     \
  }
    switch (type) {
      B2_ARITH_CASE(B2_INT8, int8_t)
      B2_ARITH_CASE(B2_UINT8, uint8_t)
      B2_ARITH_CASE(B2_INT16, int16_t)
      B2_ARITH_CASE(B2_UINT16, uint16_t)
      B2_ARITH_CASE(B2_INT32, int32_t)
      B2_ARITH_CASE(B2_UINT32, uint32_t)
      B2_ARITH_CASE(B2_INT64, int64_t)
      B2_ARITH_CASE(B2_UINT64, uint64_t)
      B2_ARITH_CASE(B2_FLOAT, float)
      B2_ARITH_CASE(B2_DOUBLE, double)
    }
#undef B2_ARITH_CASE
    if (st != B2_OK) return st;
    if (can_fail) {
      B2_RETURN_NOT_OK(slot.fetch(s));
      uint64_t row = static_cast<uint64_t>(slot.host()[0]);
      if (row != ~0ull) {
        // classify: divide-by-zero vs overflow, from the right operand of that row
        bool div = (op == B2_DIVIDE || op == B2_DIVIDE_CHECKED);
        bool zero = false;
        if (div) {
          uint64_t rv = 0;
          if (ra) {
            B2_CUDA(cudaMemcpyAsync(&rv, static_cast<const char*>(ra->data) + (ra->offset + row) * w, w,
                                    cudaMemcpyDeviceToHost, s));
            B2_CUDA(cudaStreamSynchronize(s));
          } else {
            rv = right->scalar->bits;
            if (w < 8) rv &= (1ull << (8 * w)) - 1;
          }
          if (type == B2_FLOAT) zero = (rv & 0x7fffffffull) == 0;
          else if (type == B2_DOUBLE) zero = (rv & 0x7fffffffffffffffull) == 0;
          else zero = rv == 0;
        }
        return set_error(B2_INVALID, zero ? "divide by zero" : "overflow");
      }
    }
  }
  fill_out(out, type, n, null_count, vguard.release(), data.release());
  return B2_OK;
}
