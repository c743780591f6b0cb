// render kernels for compile-time dimension 7 (mirrors the reference's tracer7 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM(kernel_set_d7, 7) }
