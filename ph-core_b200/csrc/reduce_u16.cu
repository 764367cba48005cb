// reduce_u16.cu -- reduction kernels instantiated for uint16_t (see reduce_impl.cuh).
#include "reduce_impl.cuh"
namespace ph {
template int32_t reduce_full_t<uint16_t>(int32_t, const void*, const ph_desc*, void*, int64_t*, const CombineArgs*);
template int32_t reduce_full_sharded_t<uint16_t>(int32_t, const void*, const ph_desc*, int64_t, void*, int64_t*, uint32_t*);
template int32_t reduce_axis_t<uint16_t>(int32_t, const void*, const ph_desc*, int32_t, void*, const ph_desc*);
}  // namespace ph
