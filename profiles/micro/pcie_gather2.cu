// Follow-up to pcie_gather.cu: is the ~96 M requests/s ceiling for random gathers from pinned host
// memory an address-translation limit?  Same kernel, 32-byte requests, positions confined to a
// region of 1 MiB .. 4 GiB; and the same over transparent-huge-page memory registered with
// cudaHostRegister.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <sys/mman.h>
#include <cuda_runtime.h>

__global__ void gather(const uint32_t *host, size_t n_lines, int words, int ilp, int iters, uint32_t *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t state = warp * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = 0;
            if (j < ilp) {
                state = state * 6364136223846793005ull + 1442695040888963407ull;
                const size_t pos = ((state >> 20) % n_lines) * 32;
                if (lane < words) v[j] = __ldg(host + pos + lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j];
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

static void run(const char *what, const uint32_t *dev, size_t bytes, uint32_t *sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (size_t region = 1ull << 20; region <= bytes; region <<= 2) {
        const int blocks = 148 * 8, threads = 256, iters = 200, ilp = 4;
        gather<<<blocks, threads>>>(dev, region / 128, 8, ilp, 10, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        gather<<<blocks, threads>>>(dev, region / 128, 8, ilp, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double req = (double)blocks * threads / 32 * iters * ilp;
        printf("%-22s region %6zu MiB  %8.1f Mreq/s (32 B each)\n", what, region >> 20, req / ms / 1e3);
    }
}

int main() {
    const size_t bytes = 4ull << 30;
    uint32_t *sink;
    cudaMalloc(&sink, 4);
    uint32_t *h;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = (uint32_t)i;
    run("cudaHostAlloc", h, bytes, sink);
    cudaFreeHost(h);
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(p, bytes, MADV_HUGEPAGE);
    for (size_t i = 0; i < bytes; i += 4096) ((volatile char *)p)[i] = 1;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { printf("cudaHostRegister: %s\n", cudaGetErrorString(e)); return 0; }
    void *d;
    cudaHostGetDevicePointer(&d, p, 0);
    run("THP + cudaHostRegister", (const uint32_t *)d, bytes, sink);
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) { char buf[128]; if (fgets(buf, sizeof buf, f)) printf("THP setting: %s", buf); fclose(f); }
    return 0;
}
