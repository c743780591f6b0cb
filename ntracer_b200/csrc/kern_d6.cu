// render kernels for compile-time dimension 6 (mirrors the reference's tracer6 module, fixed_geometry.hpp)
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM(kernel_set_d6, 6) }
