// ewise_f64.cu -- elementwise kernels instantiated for double (see ewise_impl.cuh).
#define PH_T double
#define PH_SUFFIX f64
#include "ewise_impl.cuh"
