// Stage 2/3 kernels for sequences of up to 512 symbols (W = 8 words); see wd_kernels23.cuh.
#include "wd_kernels23.cuh"

namespace wd {
WD_INSTANTIATE_W(8)
}  // namespace wd
