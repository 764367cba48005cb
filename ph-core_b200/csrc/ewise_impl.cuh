// ewise_impl.cuh -- per-dtype instantiation of the elementwise entry points.
// Included by ewise_<dtype>.cu with PH_T (C type) and PH_SUFFIX defined.
#pragma once
#include "map_kernels.cuh"
#include "ops.cuh"

namespace ph {

template <typename T>
inline void fill_operands2(MapOperand (&ops)[2], const void* a, const ph_desc* ad, const void* b,
                           const ph_desc* bd) {
  ops[0].base = a; ops[0].desc = ad;
  ops[1].base = b; ops[1].desc = bd;
}

// array (op) array
template <typename T>
int32_t ewise_binary_t(int32_t op, const void* a, const ph_desc* ad, const void* b, const ph_desc* bd,
                       void* out, const ph_desc* od, bool a_param, uint64_t a_bits, bool b_param,
                       uint64_t b_bits) {
  MapOperand ops[2];
  fill_operands2<T>(ops, a, ad, b, bd);
  ops[0].is_param = a_param; ops[0].param = a_bits;
  ops[1].is_param = b_param; ops[1].param = b_bits;
#define PH_CASE(OPC) case OPC: return launch_map<BinaryOp<T, OPC>>(ops, out, od);
  if constexpr (is_float_t<T>::value) {
    switch (op) {
      PH_CASE(PH_ADD) PH_CASE(PH_SUB) PH_CASE(PH_MUL) PH_CASE(PH_DIV)
      PH_CASE(PH_FLOORDIV) PH_CASE(PH_MOD) PH_CASE(PH_POW)
      case PH_POWI: return launch_map<PowiOp<T>>(ops, out, od);
      default: return set_error(PH_ERR_UNSUPPORTED, "binary op %d is not defined for float dtypes", op);
    }
  } else {
    switch (op) {
      PH_CASE(PH_ADD) PH_CASE(PH_SUB) PH_CASE(PH_MUL) PH_CASE(PH_DIV)
      PH_CASE(PH_FLOORDIV) PH_CASE(PH_MOD) PH_CASE(PH_POW)
      PH_CASE(PH_WADD) PH_CASE(PH_WSUB) PH_CASE(PH_WMUL) PH_CASE(PH_WPOW)
      PH_CASE(PH_AND) PH_CASE(PH_OR) PH_CASE(PH_XOR)
      default: return set_error(PH_ERR_UNSUPPORTED, "binary op %d is not defined for integer dtypes", op);
    }
  }
#undef PH_CASE
}

template <typename T>
int32_t compare_t(int32_t cmp, const void* a, const ph_desc* ad, const void* b, const ph_desc* bd,
                  void* out, const ph_desc* od, bool a_param, uint64_t a_bits, bool b_param,
                  uint64_t b_bits) {
  MapOperand ops[2];
  fill_operands2<T>(ops, a, ad, b, bd);
  ops[0].is_param = a_param; ops[0].param = a_bits;
  ops[1].is_param = b_param; ops[1].param = b_bits;
  switch (cmp) {
    case PH_GT: return launch_map<CompareOp<T, PH_GT>>(ops, out, od);
    case PH_LT: return launch_map<CompareOp<T, PH_LT>>(ops, out, od);
    case PH_GE: return launch_map<CompareOp<T, PH_GE>>(ops, out, od);
    case PH_LE: return launch_map<CompareOp<T, PH_LE>>(ops, out, od);
    case PH_EQ: return launch_map<CompareOp<T, PH_EQ>>(ops, out, od);
    case PH_NE: return launch_map<CompareOp<T, PH_NE>>(ops, out, od);
    default: return set_error(PH_ERR_INVALID, "unknown comparison %d", cmp);
  }
}

template <typename T>
int32_t compare3_t(const void* a, const ph_desc* ad, const void* b, const ph_desc* bd, void* out, const ph_desc* od,
                   bool a_param, uint64_t a_bits, bool b_param, uint64_t b_bits) {
  if constexpr (is_float_t<T>::value) {
    return set_error(PH_ERR_UNSUPPORTED, "<=> on floats yields Int32? (nil against NaN): not on the device path");
  } else {
    MapOperand ops[2];
    fill_operands2<T>(ops, a, ad, b, bd);
    ops[0].is_param = a_param; ops[0].param = a_bits;
    ops[1].is_param = b_param; ops[1].param = b_bits;
    return launch_map<SpaceshipOp<T>>(ops, out, od);
  }
}

template <typename T>
int32_t unary_t(int32_t op, const void* a, const ph_desc* ad, void* out, const ph_desc* od) {
  MapOperand ops[1];
  ops[0].base = a; ops[0].desc = ad;
  switch (op) {
    case PH_POS: return launch_map<UnaryOp<T, PH_POS>>(ops, out, od);
    case PH_NEG: return launch_map<UnaryOp<T, PH_NEG>>(ops, out, od);
    case PH_NOT:
      if constexpr (is_float_t<T>::value) return set_error(PH_ERR_UNSUPPORTED, "~ is not defined for floats");
      else return launch_map<UnaryOp<T, PH_NOT>>(ops, out, od);
    default: return set_error(PH_ERR_INVALID, "unknown unary op %d", op);
  }
}

template <typename T>
int32_t mul_add_t(const void* a, const ph_desc* ad, const void* b, const ph_desc* bd, const void* c,
                  const ph_desc* cd, void* out, const ph_desc* od) {
  MapOperand ops[3];
  ops[0].base = a; ops[0].desc = ad;
  ops[1].base = b; ops[1].desc = bd;
  ops[2].base = c; ops[2].desc = cd;
  return launch_map<MulAddOp<T>>(ops, out, od);
}

}  // namespace ph

#define PH_CONCAT2(a, b) a##b
#define PH_CONCAT(a, b) PH_CONCAT2(a, b)

namespace ph {
int32_t PH_CONCAT(ewise_binary_, PH_SUFFIX)(int32_t op, const void* a, const ph_desc* ad, const void* b,
                                            const ph_desc* bd, void* out, const ph_desc* od, bool ap,
                                            uint64_t abits, bool bp, uint64_t bbits) {
  return ewise_binary_t<PH_T>(op, a, ad, b, bd, out, od, ap, abits, bp, bbits);
}
int32_t PH_CONCAT(compare_, PH_SUFFIX)(int32_t cmp, const void* a, const ph_desc* ad, const void* b,
                                       const ph_desc* bd, void* out, const ph_desc* od, bool ap,
                                       uint64_t abits, bool bp, uint64_t bbits) {
  return compare_t<PH_T>(cmp, a, ad, b, bd, out, od, ap, abits, bp, bbits);
}
int32_t PH_CONCAT(compare3_, PH_SUFFIX)(const void* a, const ph_desc* ad, const void* b, const ph_desc* bd, void* out,
                                        const ph_desc* od, bool ap, uint64_t abits, bool bp, uint64_t bbits) {
  return compare3_t<PH_T>(a, ad, b, bd, out, od, ap, abits, bp, bbits);
}
int32_t PH_CONCAT(unary_, PH_SUFFIX)(int32_t op, const void* a, const ph_desc* ad, void* out,
                                     const ph_desc* od) {
  return unary_t<PH_T>(op, a, ad, out, od);
}
int32_t PH_CONCAT(mul_add_, PH_SUFFIX)(const void* a, const ph_desc* ad, const void* b, const ph_desc* bd,
                                       const void* c, const ph_desc* cd, void* out, const ph_desc* od) {
  return mul_add_t<PH_T>(a, ad, b, bd, c, cd, out, od);
}
}  // namespace ph
