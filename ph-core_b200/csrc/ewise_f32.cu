// ewise_f32.cu -- elementwise kernels instantiated for float (see ewise_impl.cuh).
#define PH_T float
#define PH_SUFFIX f32
#include "ewise_impl.cuh"
