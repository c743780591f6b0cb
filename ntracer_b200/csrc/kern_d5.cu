// render kernels for compile-time dimension 5 (mirrors the reference's tracer5 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM_WIDE(kernel_set_d5, 5) }
