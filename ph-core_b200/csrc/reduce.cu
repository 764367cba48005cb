// reduce.cu -- C-ABI entry points of the reductions (ph_reduce_full, ph_reduce_full_dev, ph_reduce_axis).
// The kernels and per-dtype launchers live in reduce_impl.cuh and are instantiated one element type per
// translation unit (reduce_<dtype>.cu); here the two launcher templates are only declared.
#include "ph_common.cuh"
#include "comm.cuh"

namespace ph {
template <typename T>
int32_t reduce_full_t(int32_t red, const void* a, const ph_desc* d, void* out_value_dev, int64_t* out_index_dev,
                      const CombineArgs* cmb);
template <typename T>
int32_t reduce_full_sharded_t(int32_t red, const void* a, const ph_desc* d, int64_t elems_before, void* out_value_host,
                              int64_t* out_index_host, uint32_t* out_flags);
template <typename T>
int32_t reduce_axis_t(int32_t red, const void* a, const ph_desc* d, int32_t axis, void* out, const ph_desc* od);
}  // namespace ph

using namespace ph;

#define PH_RED_DTYPE_SWITCH(dtype, CALL)                                                   \
  switch (dtype) {                                                                         \
    case PH_F32: return CALL(float);                                                       \
    case PH_F64: return CALL(double);                                                      \
    case PH_I32: return CALL(int32_t);                                                     \
    case PH_I64: return CALL(int64_t);                                                     \
    case PH_U8: return CALL(uint8_t);                                                      \
    case PH_I8: return CALL(int8_t);                                                       \
    case PH_I16: return CALL(int16_t);                                                     \
    case PH_U16: return CALL(uint16_t);                                                    \
    case PH_U32: return CALL(uint32_t);                                                    \
    case PH_U64: return CALL(uint64_t);                                                    \
    default: return set_error(PH_ERR_UNSUPPORTED, "dtype %d has no reduction kernels", dtype); \
  }

extern "C" {

int32_t ph_reduce_full_dev(int32_t red, int32_t dtype, const void* a, const ph_desc* a_desc,
                           void* out_value_dev, int64_t* out_index_dev) {
  PH_REQUIRE_INIT();
  if (!a || !a_desc || !out_value_dev) return set_error(PH_ERR_INVALID, "null argument to ph_reduce_full_dev");
#define CALL(T) reduce_full_t<T>(red, a, a_desc, out_value_dev, out_index_dev, nullptr)
  PH_RED_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_reduce_full(int32_t red, int32_t dtype, const void* a, const ph_desc* a_desc,
                       void* out_value_host, int64_t* out_index_host) {
  PH_REQUIRE_INIT();
  if (!out_value_host) return set_error(PH_ERR_INVALID, "null out_value_host");
  Runtime& r = rt();
  int32_t st = ensure_scratch(1 << 20);
  if (st != PH_OK) return st;
  // results live at the end of the scratch area's first MiB
  char* res = reinterpret_cast<char*>(r.d_scratch) + (1 << 20) - 64;
  st = ph_reduce_full_dev(red, dtype, a, a_desc, res, reinterpret_cast<int64_t*>(res + 16));
  if (st != PH_OK) return st;
  char* h = reinterpret_cast<char*>(r.h_scratch) + 64;
  PH_CUDA(cudaMemcpyAsync(h, res, 32, cudaMemcpyDeviceToHost, r.stream));
  PH_CUDA(cudaStreamSynchronize(r.stream));
  memcpy(out_value_host, h, dtype_size(dtype));
  if (out_index_host) memcpy(out_index_host, h + 16, 8);
  return PH_OK;
}

int32_t ph_reduce_full_sharded(int32_t red, int32_t dtype, const void* a, const ph_desc* a_desc, int64_t elems_before,
                               void* out_value_host, int64_t* out_index_host, uint32_t* out_flags) {
  PH_REQUIRE_INIT();
  if (!a || !a_desc || !out_value_host) return set_error(PH_ERR_INVALID, "null argument to ph_reduce_full_sharded");
  if (red < PH_SUM || red > PH_ARGMIN) return set_error(PH_ERR_INVALID, "unknown reduction %d", red);
#define CALL(T) reduce_full_sharded_t<T>(red, a, a_desc, elems_before, out_value_host, out_index_host, out_flags)
  PH_RED_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

int32_t ph_reduce_axis(int32_t red, int32_t dtype, const void* a, const ph_desc* a_desc, int32_t axis,
                       void* out, const ph_desc* out_desc) {
  PH_REQUIRE_INIT();
  if (!a || !a_desc || !out || !out_desc) return set_error(PH_ERR_INVALID, "null argument to ph_reduce_axis");
#define CALL(T) reduce_axis_t<T>(red, a, a_desc, axis, out, out_desc)
  PH_RED_DTYPE_SWITCH(dtype, CALL)
#undef CALL
}

}  // extern "C"
