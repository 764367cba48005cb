// ewise_u16.cu -- elementwise kernels instantiated for uint16_t (see ewise_impl.cuh).
#define PH_T uint16_t
#define PH_SUFFIX u16
#include "ewise_impl.cuh"
