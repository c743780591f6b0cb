// render kernels for compile-time dimension 8 (mirrors the reference's tracer8 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM(kernel_set_d8, 8) }
