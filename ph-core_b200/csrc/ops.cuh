// ops.cuh -- element operators with Crystal 1.0.0 number semantics.
// The operator list is src/multi_indexable.cr:960-985.  Floats: every operator is
// one IEEE round-to-nearest operation (the __f*_rn / __d*_rn intrinsics are never
// contracted into FMAs; the library is also built with -fmad=false).  Ints:
// + - * and unary - are overflow-checked, &+ &- &* wrap, // and % are floored.
#pragma once
#include "ph_common.cuh"

namespace ph {

template <typename T> struct is_float_t : std::false_type {};
template <> struct is_float_t<float> : std::true_type {};
template <> struct is_float_t<double> : std::true_type {};

template <typename T> struct int_limits;
template <> struct int_limits<int8_t>  { static constexpr int64_t lo = INT8_MIN,  hi = INT8_MAX; };
template <> struct int_limits<int16_t> { static constexpr int64_t lo = INT16_MIN, hi = INT16_MAX; };
template <> struct int_limits<int32_t> { static constexpr int64_t lo = INT32_MIN, hi = INT32_MAX; };
template <> struct int_limits<uint8_t>  { static constexpr int64_t lo = 0, hi = UINT8_MAX; };
template <> struct int_limits<uint16_t> { static constexpr int64_t lo = 0, hi = UINT16_MAX; };
template <> struct int_limits<uint32_t> { static constexpr int64_t lo = 0, hi = UINT32_MAX; };

// ---- float primitives (single rounding each)
__device__ __forceinline__ float  f_add(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ float  f_sub(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ float  f_mul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ float  f_div(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double f_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double f_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double f_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double f_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  f_floor(float a)  { return floorf(a); }
__device__ __forceinline__ double f_floor(double a) { return floor(a); }

// ---- checked / wrapping integer primitives
template <typename T>
__device__ __forceinline__ T wrap_from_i64(int64_t v) { return (T)v; }   // two's complement truncation

template <typename T>
__device__ __forceinline__ T i_add(T a, T b, bool checked, uint32_t& err) {
  if constexpr (sizeof(T) < 8) {
    const int64_t r = (int64_t)a + (int64_t)b;
    if (checked && (r < int_limits<T>::lo || r > int_limits<T>::hi)) err |= PH_FLAG_OVERFLOW;
    return wrap_from_i64<T>(r);
  } else if constexpr (std::is_signed<T>::value) {
    const uint64_t r = (uint64_t)a + (uint64_t)b;
    if (checked && (int64_t)(((uint64_t)a ^ r) & ((uint64_t)b ^ r)) < 0) err |= PH_FLAG_OVERFLOW;
    return (T)r;
  } else {
    const uint64_t r = (uint64_t)a + (uint64_t)b;
    if (checked && r < (uint64_t)a) err |= PH_FLAG_OVERFLOW;
    return (T)r;
  }
}
template <typename T>
__device__ __forceinline__ T i_sub(T a, T b, bool checked, uint32_t& err) {
  if constexpr (sizeof(T) < 8) {
    const int64_t r = (int64_t)a - (int64_t)b;
    if (checked && (r < int_limits<T>::lo || r > int_limits<T>::hi)) err |= PH_FLAG_OVERFLOW;
    return wrap_from_i64<T>(r);
  } else if constexpr (std::is_signed<T>::value) {
    const uint64_t r = (uint64_t)a - (uint64_t)b;
    if (checked && (int64_t)(((uint64_t)a ^ (uint64_t)b) & ((uint64_t)a ^ r)) < 0) err |= PH_FLAG_OVERFLOW;
    return (T)r;
  } else {
    if (checked && (uint64_t)b > (uint64_t)a) err |= PH_FLAG_OVERFLOW;
    return (T)((uint64_t)a - (uint64_t)b);
  }
}
template <typename T>
__device__ __forceinline__ T i_mul(T a, T b, bool checked, uint32_t& err) {
  if constexpr (sizeof(T) < 8) {
    const int64_t r = (int64_t)a * (int64_t)b;
    if (checked && (r < int_limits<T>::lo || r > int_limits<T>::hi)) err |= PH_FLAG_OVERFLOW;
    return wrap_from_i64<T>(r);
  } else if constexpr (std::is_signed<T>::value) {
    const int64_t lo = (int64_t)((uint64_t)a * (uint64_t)b);
    const int64_t hi = __mul64hi((int64_t)a, (int64_t)b);
    if (checked && hi != (lo >> 63)) err |= PH_FLAG_OVERFLOW;
    return (T)lo;
  } else {
    if (checked && __umul64hi((uint64_t)a, (uint64_t)b) != 0) err |= PH_FLAG_OVERFLOW;
    return (T)((uint64_t)a * (uint64_t)b);
  }
}
// Int#// : floored; /0 -> DivisionByZeroError; MIN // -1 -> ArgumentError
template <typename T>
__device__ __forceinline__ T i_floordiv(T a, T b, uint32_t& err) {
  if (b == 0) { err |= PH_FLAG_DIV0; return 0; }
  if constexpr (std::is_signed<T>::value) {
    if (b == (T)-1) {
      if (a == (T)((uint64_t)1 << (sizeof(T) * 8 - 1))) { err |= PH_FLAG_ARGUMENT; return a; }
      return (T)(-a);
    }
    T q = a / b;
    const T m = a - q * b;
    if (m != 0 && ((m < 0) != (b < 0))) q -= 1;
    return q;
  } else {
    return a / b;
  }
}
// Int#% : floored modulo (sign of the divisor)
template <typename T>
__device__ __forceinline__ T i_mod(T a, T b, uint32_t& err) {
  if (b == 0) { err |= PH_FLAG_DIV0; return 0; }
  if constexpr (std::is_signed<T>::value) {
    if (b == (T)-1) return 0;
    T m = a % b;
    if (m != 0 && ((m < 0) != (b < 0))) m += b;
    return m;
  } else {
    return a % b;
  }
}
// Int#** / Int#&** : square-and-multiply, `k *= k` only while exponent bits remain
template <typename T>
__device__ __forceinline__ T i_pow(T base, T exp, bool checked, uint32_t& err) {
  if constexpr (std::is_signed<T>::value) {
    if (exp < 0) { err |= PH_FLAG_ARGUMENT; return 0; }
  }
  T result = 1;
  T k = base;
  typename std::make_unsigned<T>::type e = (typename std::make_unsigned<T>::type)exp;
  while (e > 0) {
    if (e & 1) result = i_mul<T>(result, k, checked, err);
    e >>= 1;
    if (e > 0) k = i_mul<T>(k, k, checked, err);
  }
  return result;
}
// Float ** Int32 : llvm.powi -> compiler-rt __powisf2 / __powidf2
template <typename T>
__device__ __forceinline__ T f_powi(T a, int32_t n) {
  const bool recip = n < 0;
  T r = (T)1;
  int32_t b = n;
  while (true) {
    if (b & 1) r = f_mul(r, a);
    b /= 2;
    if (b == 0) break;
    a = f_mul(a, a);
  }
  return recip ? f_div((T)1, r) : r;
}

// Hot (op, dtype) pairs get every vector-width variant of the map kernels; the rest only the
// widest and the scalar one (keeps the library small; alignment-odd views of rare ops still work).
template <typename T> struct is_main_t : std::false_type {};
template <> struct is_main_t<float> : std::true_type {};
template <> struct is_main_t<double> : std::true_type {};
template <> struct is_main_t<int32_t> : std::true_type {};
template <> struct is_main_t<int64_t> : std::true_type {};

// ---------------------------------------------------------------- binary functors
template <typename T, int OP> struct BinOut { using type = T; };
template <> struct BinOut<int32_t, PH_DIV> { using type = double; };
template <> struct BinOut<int64_t, PH_DIV> { using type = double; };
template <> struct BinOut<uint8_t, PH_DIV> { using type = double; };
template <> struct BinOut<int8_t, PH_DIV> { using type = double; };
template <> struct BinOut<int16_t, PH_DIV> { using type = double; };
template <> struct BinOut<uint16_t, PH_DIV> { using type = double; };
template <> struct BinOut<uint32_t, PH_DIV> { using type = double; };
template <> struct BinOut<uint64_t, PH_DIV> { using type = double; };

template <typename T, int OP>
struct BinaryOp {
  using In = T;
  using Out = typename BinOut<T, OP>::type;
  static constexpr int NIN = 2;
  static constexpr bool kCompact = !(is_main_t<T>::value && (OP == PH_ADD || OP == PH_SUB || OP == PH_MUL || OP == PH_DIV));
  static __device__ __forceinline__ Out apply(const In (&x)[2], uint32_t& err) {
    const T a = x[0], b = x[1];
    if constexpr (is_float_t<T>::value) {
      if constexpr (OP == PH_ADD || OP == PH_WADD) return f_add(a, b);
      else if constexpr (OP == PH_SUB || OP == PH_WSUB) return f_sub(a, b);
      else if constexpr (OP == PH_MUL || OP == PH_WMUL) return f_mul(a, b);
      else if constexpr (OP == PH_DIV) return f_div(a, b);
      else if constexpr (OP == PH_FLOORDIV) return f_floor(f_div(a, b));
      else if constexpr (OP == PH_MOD) {
        if (b == (T)0) err |= PH_FLAG_DIV0;
        return f_sub(a, f_mul(b, f_floor(f_div(a, b))));
      } else if constexpr (OP == PH_POW) {       // libm pow: tolerance-only parity
        if constexpr (sizeof(T) == 4) return powf(a, b);
        else return pow(a, b);
      } else return a;
    } else {
      if constexpr (OP == PH_ADD) return i_add<T>(a, b, true, err);
      else if constexpr (OP == PH_SUB) return i_sub<T>(a, b, true, err);
      else if constexpr (OP == PH_MUL) return i_mul<T>(a, b, true, err);
      else if constexpr (OP == PH_WADD) return i_add<T>(a, b, false, err);
      else if constexpr (OP == PH_WSUB) return i_sub<T>(a, b, false, err);
      else if constexpr (OP == PH_WMUL) return i_mul<T>(a, b, false, err);
      else if constexpr (OP == PH_DIV) return __ddiv_rn((double)a, (double)b);
      else if constexpr (OP == PH_FLOORDIV) return i_floordiv<T>(a, b, err);
      else if constexpr (OP == PH_MOD) return i_mod<T>(a, b, err);
      else if constexpr (OP == PH_POW) return i_pow<T>(a, b, true, err);
      else if constexpr (OP == PH_WPOW) return i_pow<T>(a, b, false, err);
      else if constexpr (OP == PH_AND) return a & b;
      else if constexpr (OP == PH_OR) return a | b;
      else if constexpr (OP == PH_XOR) return a ^ b;
      else return a;
    }
  }
};

// Float ** Int32 scalar.  x[1] carries the exponent's bits reinterpreted in T's width.
template <typename T>
struct PowiOp {
  using In = T;
  using Out = T;
  static constexpr int NIN = 2;
  static constexpr bool kCompact = true;
  static __device__ __forceinline__ Out apply(const In (&x)[2], uint32_t&) {
    int32_t n;
    if constexpr (sizeof(T) == 4) n = __float_as_int(x[1]);
    else n = (int32_t)(__double_as_longlong(x[1]) & 0xffffffffLL);
    return f_powi<T>(x[0], n);
  }
};

template <typename T, int CMP>
struct CompareOp {
  using In = T;
  using Out = uint8_t;
  static constexpr int NIN = 2;
  static constexpr bool kCompact = !is_main_t<T>::value;
  static __device__ __forceinline__ Out apply(const In (&x)[2], uint32_t&) {
    if constexpr (CMP == PH_GT) return x[0] > x[1];
    else if constexpr (CMP == PH_LT) return x[0] < x[1];
    else if constexpr (CMP == PH_GE) return x[0] >= x[1];
    else if constexpr (CMP == PH_LE) return x[0] <= x[1];
    else if constexpr (CMP == PH_EQ) return x[0] == x[1];
    else return x[0] != x[1];
  }
};

// `<=>` (src/multi_indexable.cr:981 in the operator list): Int#<=> yields -1 / 0 / 1 as Int32.  Integer
// element types only: Float#<=> is Int32? (nil against NaN), which has no device representation.
template <typename T>
struct SpaceshipOp {
  using In = T;
  using Out = int32_t;
  static constexpr int NIN = 2;
  static constexpr bool kCompact = true;
  static __device__ __forceinline__ Out apply(const In (&x)[2], uint32_t&) {
    return (x[0] > x[1]) - (x[0] < x[1]);
  }
};

template <typename T, int OP>
struct UnaryOp {
  using In = T;
  using Out = T;
  static constexpr int NIN = 1;
  static constexpr bool kCompact = true;
  static __device__ __forceinline__ Out apply(const In (&x)[1], uint32_t& err) {
    if constexpr (OP == PH_POS) return x[0];
    else if constexpr (OP == PH_NEG) {
      if constexpr (is_float_t<T>::value) return -x[0];
      else return i_sub<T>((T)0, x[0], true, err);
    } else {
      if constexpr (is_float_t<T>::value) return x[0];
      else return (T)~x[0];
    }
  }
};

// out = (a * b) + c, two roundings (reference: `a * b` materialises, then `+ c`)
template <typename T>
struct MulAddOp {
  using In = T;
  using Out = T;
  static constexpr int NIN = 3;
  static constexpr bool kCompact = !is_main_t<T>::value;
  static __device__ __forceinline__ Out apply(const In (&x)[3], uint32_t& err) {
    if constexpr (is_float_t<T>::value) return f_add(f_mul(x[0], x[1]), x[2]);
    else return i_add<T>(i_mul<T>(x[0], x[1], true, err), x[2], true, err);
  }
};

}  // namespace ph
