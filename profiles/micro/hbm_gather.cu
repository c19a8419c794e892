// How many scattered 32-byte sectors per second can a B200 pull from HBM?
// (the ceiling of the fused counting kernel, whose reads are 1 byte per 32-byte sector touched)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hbm_gather hbm_gather.cu && ./hbm_gather
// Each thread issues U independent 1-byte loads per iteration at pseudo-random sector addresses of
// a buffer far larger than L2; the pattern "clustered" mimics a target: 12 sectors 1571 bytes apart.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int U, int CLUSTER>
__global__ void gather(const uint8_t *buf, uint64_t n_sectors, uint64_t loads_per_thread, uint32_t *out, uint64_t seed) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t it = 0; it < loads_per_thread; it += U) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t sec;
            if (CLUSTER == 1) {
                sec = mix(seed + (tid * loads_per_thread + it + u)) % n_sectors;
            } else {
                // CLUSTER consecutive loads of a thread walk rows 1571 bytes apart from a random origin
                const uint64_t g = (it + u) / CLUSTER, r = (it + u) % CLUSTER;
                const uint64_t origin = mix(seed + tid * loads_per_thread + g) % (n_sectors - 4096);
                sec = origin + (r * 1571) / 32;
            }
            v[u] = __ldg(buf + sec * 32 + (tid & 31));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

template <int U, int CLUSTER>
static void run(const char *name, const uint8_t *buf, uint64_t n_sectors, uint32_t *out, int blocks, int threads) {
    const uint64_t per_thread = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<U, CLUSTER><<<blocks, threads>>>(buf, n_sectors, per_thread, out, 1);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) gather<U, CLUSTER><<<blocks, threads>>>(buf, n_sectors, per_thread, out, 100 + r);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double loads = 5.0 * blocks * threads * per_thread;
    printf("%-34s U=%2d blocks=%5d x %4d: %7.1f G sector loads/s = %6.2f TB/s of 32-B sectors (%.3f ms per launch)\n", name, U, blocks,
           threads, loads / (ms * 1e-3) / 1e9, loads * 32 / (ms * 1e-3) / 1e12, ms / 5);
}

int main() {
    const uint64_t bytes = 16ull << 30;
    uint8_t *buf;
    uint32_t *out;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4);
    cudaMemset(buf, 1, bytes);
    const uint64_t n_sectors = bytes / 32;
    for (int threads : {256, 1024}) {
        const int blocks = 148 * (2048 / threads) * 4;
        run<1, 1>("random sectors", buf, n_sectors, out, blocks, threads);
        run<4, 1>("random sectors", buf, n_sectors, out, blocks, threads);
        run<8, 1>("random sectors", buf, n_sectors, out, blocks, threads);
        run<16, 1>("random sectors", buf, n_sectors, out, blocks, threads);
        run<8, 12>("12-row clusters (1571 B apart)", buf, n_sectors, out, blocks, threads);
        run<16, 12>("12-row clusters (1571 B apart)", buf, n_sectors, out, blocks, threads);
    }
    // every lane of a warp its own sector (above: the 32 lanes of a warp share a sector? no: sec depends on tid) -- note
    return 0;
}
