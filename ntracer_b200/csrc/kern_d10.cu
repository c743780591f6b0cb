// render kernels for compile-time dimension 10.  The reference builds fixed-dimension modules for 3..8 by default
// (setup.py --optimize-dimensions); 10 is added here because BASELINE config 5 / the 9-D soup fixture live there.
#include "kernels.cuh"
namespace ntr { NTR_INSTANTIATE_DIM(kernel_set_d10, 10) }
