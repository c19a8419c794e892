// Stage 2/3 kernels for sequences of up to 256 symbols (W = 4 words); see wd_kernels23.cuh.
#include "wd_kernels23.cuh"

namespace wd {
WD_INSTANTIATE_W(4)
}  // namespace wd
