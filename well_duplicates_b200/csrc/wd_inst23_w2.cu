// Stage 2/3 kernels for sequences of up to 128 symbols (W = 2 words); see wd_kernels23.cuh.
#include "wd_kernels23.cuh"

namespace wd {
WD_INSTANTIATE_W(2)
}  // namespace wd
