// ewise_i64.cu -- elementwise kernels instantiated for int64_t (see ewise_impl.cuh).
#define PH_T int64_t
#define PH_SUFFIX i64
#include "ewise_impl.cuh"
