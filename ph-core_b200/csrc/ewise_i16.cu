// ewise_i16.cu -- elementwise kernels instantiated for int16_t (see ewise_impl.cuh).
#define PH_T int16_t
#define PH_SUFFIX i16
#include "ewise_impl.cuh"
