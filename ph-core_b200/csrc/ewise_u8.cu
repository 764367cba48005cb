// ewise_u8.cu -- elementwise kernels instantiated for uint8_t (see ewise_impl.cuh).
#define PH_T uint8_t
#define PH_SUFFIX u8
#include "ewise_impl.cuh"
