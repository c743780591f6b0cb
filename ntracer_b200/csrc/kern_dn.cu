// render kernels with a run-time dimension (mirrors the reference's generic tracern module, var_geometry.hpp)
#define NTR_GENERIC_UNIT 1
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM(kernel_set_dn, 0) }
