// ewise_u64.cu -- elementwise kernels instantiated for uint64_t (see ewise_impl.cuh).
#define PH_T uint64_t
#define PH_SUFFIX u64
#include "ewise_impl.cuh"
