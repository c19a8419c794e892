// Stage 2/3 kernels for sequences of up to 64 symbols (W = 1 words); see wd_kernels23.cuh.
#include "wd_kernels23.cuh"

namespace wd {
WD_INSTANTIATE_W(1)
}  // namespace wd
