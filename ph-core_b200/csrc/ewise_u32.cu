// ewise_u32.cu -- elementwise kernels instantiated for uint32_t (see ewise_impl.cuh).
#define PH_T uint32_t
#define PH_SUFFIX u32
#include "ewise_impl.cuh"
