// render kernels for compile-time dimension 3 (mirrors the reference's tracer3 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM_WIDE(kernel_set_d3, 3) }
