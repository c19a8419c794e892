// Stage 2/3 kernels for sequences of up to 1024 symbols (W = 16 words); see wd_kernels23.cuh.
#include "wd_kernels23.cuh"

namespace wd {
WD_INSTANTIATE_W(16)
}  // namespace wd
