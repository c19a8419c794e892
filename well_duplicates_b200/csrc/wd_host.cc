// Host-memory helpers of the C ABI that involve no kernels: page-locking memory the caller already owns, so that
// ordinary buffers (a numpy array, an mmap-ed run folder cache) can be handed to wd_tile_map_host like blocks from
// wd_host_alloc -- the reference's callers hold their planes in plain Python bytes (bcl_direct_reader.py:344-345).
#include "wd_common.cuh"

using namespace wd;

extern "C" {

int wd_host_register(void *p, size_t bytes) {
    if (p == nullptr || bytes == 0) WD_FAIL(WD_E_ARG, "wd_host_register: null or empty range");
    WD_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return WD_OK;
}

int wd_host_unregister(void *p) {
    if (p == nullptr) return WD_OK;
    WD_CUDA(cudaHostUnregister(p));
    return WD_OK;
}

}  // extern "C"
