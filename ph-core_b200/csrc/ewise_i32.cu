// ewise_i32.cu -- elementwise kernels instantiated for int32_t (see ewise_impl.cuh).
#define PH_T int32_t
#define PH_SUFFIX i32
#include "ewise_impl.cuh"
