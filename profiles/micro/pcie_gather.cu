// Microbenchmark: how fast can a B200 kernel pull scattered pieces of pinned host memory across PCIe?
// Each warp reads `span` contiguous bytes (32, 64, 128: lanes read adjacent 4-byte words) at a random
// span-aligned position of a 4 GiB pinned host buffer; `ilp` independent positions are in flight per warp.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_gather pcie_gather.cu && ./pcie_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void gather(const uint32_t *host, size_t n_words, int span, int ilp, int iters, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int words = span / 4;                     // lanes < words take part
    uint64_t state = warp * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = 0;
            if (j < ilp) {
                state = state * 6364136223846793005ull + 1442695040888963407ull;
                const size_t pos = ((state >> 20) % (n_words / 32)) * 32;   // 128-byte aligned
                if (lane < words) v[j] = __ldg(host + pos + lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

int main() {
    const size_t bytes = 4ull << 30;
    uint32_t *h, *sink;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = (uint32_t)i;
    cudaMalloc(&sink, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    printf("span ilp warps/SM   Mreq/s   GB/s(payload)\n");
    for (int span : {32, 64, 128}) {
        for (int ilp : {1, 4, 8}) {
            for (int cta_per_sm : {2, 8}) {
                const int blocks = 148 * cta_per_sm, threads = 256, iters = 200;
                gather<<<blocks, threads>>>(h, bytes / 4, span, ilp, 20, sink);
                cudaDeviceSynchronize();
                cudaEventRecord(a);
                gather<<<blocks, threads>>>(h, bytes / 4, span, ilp, iters, sink);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                const double req = (double)blocks * threads / 32 * iters * ilp;
                printf("%4d %3d %6d   %8.1f   %8.2f\n", span, ilp, cta_per_sm * 8, req / ms / 1e3, req * span / ms / 1e6);
            }
        }
    }
    return 0;
}
