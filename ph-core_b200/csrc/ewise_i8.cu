// ewise_i8.cu -- elementwise kernels instantiated for int8_t (see ewise_impl.cuh).
#define PH_T int8_t
#define PH_SUFFIX i8
#include "ewise_impl.cuh"
