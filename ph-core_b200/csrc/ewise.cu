// ewise.cu -- C-ABI entry points of the elementwise family; dispatch on dtype.
// Replaces def_elementwise_binary (src/multi_indexable.cr:931-952), the scalar
// overloads (:947-951, src/patches/number.cr:6-15), unary ops (:954-958) and the
// comparison ops (:977-980, eq :899-920).
#include "ph_common.cuh"

namespace ph {
#define PH_DECL(S)                                                                                         \
  int32_t ewise_binary_##S(int32_t, const void*, const ph_desc*, const void*, const ph_desc*, void*,      \
                           const ph_desc*, bool, uint64_t, bool, uint64_t);                                \
  int32_t compare_##S(int32_t, const void*, const ph_desc*, const void*, const ph_desc*, void*,           \
                      const ph_desc*, bool, uint64_t, bool, uint64_t);                                     \
  int32_t compare3_##S(const void*, const ph_desc*, const void*, const ph_desc*, void*, const ph_desc*,   \
                       bool, uint64_t, bool, uint64_t);                                                    \
  int32_t unary_##S(int32_t, const void*, const ph_desc*, void*, const ph_desc*);                          \
  int32_t mul_add_##S(const void*, const ph_desc*, const void*, const ph_desc*, const void*,              \
                      const ph_desc*, void*, const ph_desc*);
PH_DECL(f32) PH_DECL(f64) PH_DECL(i32) PH_DECL(i64)
PH_DECL(u8) PH_DECL(i8) PH_DECL(i16) PH_DECL(u16) PH_DECL(u32) PH_DECL(u64)
#undef PH_DECL
}  // namespace ph

using namespace ph;

#define PH_DTYPE_SWITCH(dtype, CALL)                                                  \
  switch (dtype) {                                                                    \
    case PH_F32: return CALL(f32);                                                    \
    case PH_F64: return CALL(f64);                                                    \
    case PH_I32: return CALL(i32);                                                    \
    case PH_I64: return CALL(i64);                                                    \
    case PH_U8: return CALL(u8);                                                      \
    case PH_I8: return CALL(i8);                                                      \
    case PH_I16: return CALL(i16);                                                    \
    case PH_U16: return CALL(u16);                                                    \
    case PH_U32: return CALL(u32);                                                    \
    case PH_U64: return CALL(u64);                                                    \
    default: return set_error(PH_ERR_UNSUPPORTED, "dtype %d has no arithmetic kernels", dtype); \
  }

static uint64_t scalar_bits_for(int32_t op, int32_t dtype, const void* scalar_host) {
  if (op == PH_POWI) {            // int32 exponent carried in the low 32 bits
    uint32_t n;
    memcpy(&n, scalar_host, 4);
    return (uint64_t)n;
  }
  return host_scalar_bits(scalar_host, dtype_size(dtype));
}

extern "C" {

int32_t ph_ewise_binary(int32_t op, int32_t dtype, const void* a, const ph_desc* a_desc, const void* b,
                        const ph_desc* b_desc, void* out, const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
  if (op == PH_POWI) return set_error(PH_ERR_INVALID, "PH_POWI takes a scalar exponent (ph_ewise_scalar)");
#define CALL(S) ewise_binary_##S(op, a, a_desc, b, b_desc, out, out_desc, false, 0, false, 0)
  PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_ewise_scalar(int32_t op, int32_t dtype, const void* a, const ph_desc* a_desc,
                        const void* scalar_host, int32_t scalar_on_left, void* out,
                        const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
  if (!scalar_host) return set_error(PH_ERR_INVALID, "null scalar");
  if (op == PH_POWI && scalar_on_left) return set_error(PH_ERR_INVALID, "PH_POWI: the exponent is the right operand");
  const uint64_t bits = scalar_bits_for(op, dtype, scalar_host);
  if (scalar_on_left) {
#define CALL(S) ewise_binary_##S(op, nullptr, nullptr, a, a_desc, out, out_desc, true, bits, false, 0)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  } else {
#define CALL(S) ewise_binary_##S(op, a, a_desc, nullptr, nullptr, out, out_desc, false, 0, true, bits)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  }
}

int32_t ph_ewise_unary(int32_t op, int32_t dtype, const void* a, const ph_desc* a_desc, void* out,
                       const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
#define CALL(S) unary_##S(op, a, a_desc, out, out_desc)
  PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_ewise_mul_add(int32_t dtype, const void* a, const ph_desc* a_desc, const void* b,
                         const ph_desc* b_desc, const void* c, const ph_desc* c_desc, void* out,
                         const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
#define CALL(S) mul_add_##S(a, a_desc, b, b_desc, c, c_desc, out, out_desc)
  PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_compare(int32_t cmp, int32_t dtype, const void* a, const ph_desc* a_desc, const void* b,
                   const ph_desc* b_desc, uint8_t* out, const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
#define CALL(S) compare_##S(cmp, a, a_desc, b, b_desc, out, out_desc, false, 0, false, 0)
  PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_compare_scalar(int32_t cmp, int32_t dtype, const void* a, const ph_desc* a_desc,
                          const void* scalar_host, int32_t scalar_on_left, uint8_t* out,
                          const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
  if (!scalar_host) return set_error(PH_ERR_INVALID, "null scalar");
  const uint64_t bits = host_scalar_bits(scalar_host, dtype_size(dtype));
  if (scalar_on_left) {
#define CALL(S) compare_##S(cmp, nullptr, nullptr, a, a_desc, out, out_desc, true, bits, false, 0)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  } else {
#define CALL(S) compare_##S(cmp, a, a_desc, nullptr, nullptr, out, out_desc, false, 0, true, bits)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  }
}

int32_t ph_compare3(int32_t dtype, const void* a, const ph_desc* a_desc, const void* b, const ph_desc* b_desc,
                    int32_t* out, const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
#define CALL(S) compare3_##S(a, a_desc, b, b_desc, out, out_desc, false, 0, false, 0)
  PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_compare3_scalar(int32_t dtype, const void* a, const ph_desc* a_desc, const void* scalar_host,
                           int32_t scalar_on_left, int32_t* out, const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
  if (!scalar_host) return set_error(PH_ERR_INVALID, "null scalar");
  const uint64_t bits = host_scalar_bits(scalar_host, dtype_size(dtype));
  if (scalar_on_left) {
#define CALL(S) compare3_##S(nullptr, nullptr, a, a_desc, out, out_desc, true, bits, false, 0)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  } else {
#define CALL(S) compare3_##S(a, a_desc, nullptr, nullptr, out, out_desc, false, 0, true, bits)
    PH_DTYPE_SWITCH(dtype, CALL)
#undef CALL
  }
}

}  // extern "C"
