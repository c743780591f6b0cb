// render kernels for compile-time dimension 4 (mirrors the reference's tracer4 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM_WIDE(kernel_set_d4, 4) }
