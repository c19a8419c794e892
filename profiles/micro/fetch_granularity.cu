// What does one scattered 1-byte load cost in DRAM traffic on a B200, and does any load flavour change it?
// (the fused counting kernel reads a few bytes per 32-byte sector; ncu showed ~1.9 DRAM sectors per requested one)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fetch_granularity fetch_granularity.cu
//   ./fetch_granularity [l2_fetch_limit]     # run under: ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
// Every kernel issues the same number of loads at pseudo-random 32-byte sectors of an 8 GB buffer (>> L2);
// kernels differ only in the load instruction.  The "span" kernels read 2 or 4 neighbouring sectors of one
// 64-/128-byte block from neighbouring lanes: traffic per block tells the fill granularity.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

enum Flavour { NC, CG, CV, NC_NOALLOC, EVICT_FIRST, NC_L2_64, NC_L2_128, NC_L2_256, V4, CP_ASYNC4, CP_ASYNC16, BULK16, LU,
               NC_EVICT_FIRST_POLICY };

template <int F>
__device__ __forceinline__ uint32_t load1(const uint8_t *p, uint32_t *smem_slot, uint64_t policy) {
    uint32_t v = 0;
    if (F == NC) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == CG) asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == CV) asm volatile("ld.global.cv.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == LU) asm volatile("ld.global.lu.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == EVICT_FIRST) asm volatile("ld.global.L1::evict_first.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == NC_L2_64) asm volatile("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == NC_L2_128) asm volatile("ld.global.nc.L2::128B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == NC_L2_256) asm volatile("ld.global.nc.L2::256B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    if (F == NC_EVICT_FIRST_POLICY)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    if (F == V4) {
        uint32_t a, b, c, d;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"((const void *)((uintptr_t)p & ~(uintptr_t)15)));
        v = a ^ b ^ c ^ d;
    }
    if (F == CP_ASYNC4) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_slot);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"((const void *)((uintptr_t)p & ~(uintptr_t)3)));
    }
    if (F == CP_ASYNC16) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_slot);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"((const void *)((uintptr_t)p & ~(uintptr_t)15)));
    }
    return v;
}

// SPAN = sectors of one aligned block read by SPAN neighbouring lanes (1: every lane its own random sector)
template <int F, int SPAN, int U>
__global__ void __launch_bounds__(256) probe(const uint8_t *buf, uint64_t n_sectors, int iters, uint32_t *out, uint64_t seed) {
    __shared__ __align__(16) uint32_t smem[256 * 4 * U];
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t policy = 0;
    if (F == NC_EVICT_FIRST_POLICY) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t grp = (tid / SPAN) * (uint64_t)iters * U + (uint64_t)it * U + u;
            const uint64_t sec = (mix(seed + grp) % (n_sectors / SPAN)) * SPAN + (tid % SPAN);
            v[u] = load1<F>(buf + sec * 32 + (tid & 15), &smem[(threadIdx.x * U + u) * 4], policy);
        }
        if (F == CP_ASYNC4 || F == CP_ASYNC16) {
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = smem[(threadIdx.x * U + u) * 4];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

// one elected lane per warp pulls 16 bytes per request with a bulk copy (TMA unit), 32 requests in flight per warp
__global__ void __launch_bounds__(256) probe_bulk16(const uint8_t *buf, uint64_t n_sectors, int iters, uint32_t *out, uint64_t seed) {
    __shared__ __align__(16) uint32_t smem[256 * 4];
    __shared__ __align__(8) uint64_t bar[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    __syncwarp();
    uint32_t acc = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
        const uint64_t sec = mix(seed + tid * (uint64_t)iters + it) % n_sectors;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&smem[threadIdx.x * 4]);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32 * 16));
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst),
                     "l"(buf + sec * 32), "r"(b)
                     : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(phase) : "memory");
        phase ^= 1;
        acc += smem[threadIdx.x * 4];
        __syncwarp();
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

template <int F, int SPAN, int U>
static void run(const char *name, const uint8_t *buf, uint64_t n_sectors, uint32_t *out) {
    const int blocks = 148 * 8 * 4, threads = 256, iters = 64 / U;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<F, SPAN, U><<<blocks, threads>>>(buf, n_sectors, iters, out, 7);
    cudaEventRecord(a);
    probe<F, SPAN, U><<<blocks, threads>>>(buf, n_sectors, iters, out, 100);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double loads = (double)blocks * threads * iters * U;
    printf("%-44s span=%d U=%d: %6.1f G loads/s, %.3f ms, %.0f loads (%s)\n", name, SPAN, U, loads / (ms * 1e-3) / 1e9, ms, loads,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity <- %s: %s, now %zu\n", argv[1], cudaGetErrorString(e), got);
    }
    const uint64_t bytes = 8ull << 30;
    uint8_t *buf;
    uint32_t *out;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4);
    cudaMemset(buf, 1, bytes);
    const uint64_t n_sectors = bytes / 32;
    run<NC, 1, 8>("ld.global.nc.u8", buf, n_sectors, out);
    run<CG, 1, 8>("ld.global.cg.u8", buf, n_sectors, out);
    run<CV, 1, 8>("ld.global.cv.u8", buf, n_sectors, out);
    run<LU, 1, 8>("ld.global.lu.u8", buf, n_sectors, out);
    run<NC_NOALLOC, 1, 8>("ld.global.nc.L1::no_allocate.u8", buf, n_sectors, out);
    run<EVICT_FIRST, 1, 8>("ld.global.L1::evict_first.u8", buf, n_sectors, out);
    run<NC_EVICT_FIRST_POLICY, 1, 8>("ld.global.nc + L2 evict_first policy", buf, n_sectors, out);
    run<NC_L2_64, 1, 8>("ld.global.nc.L2::64B.u8", buf, n_sectors, out);
    run<NC_L2_128, 1, 8>("ld.global.nc.L2::128B.u8", buf, n_sectors, out);
    run<NC_L2_256, 1, 8>("ld.global.nc.L2::256B.u8", buf, n_sectors, out);
    run<V4, 1, 8>("ld.global.nc.v4.u32", buf, n_sectors, out);
    run<CP_ASYNC4, 1, 8>("cp.async.ca 4 B", buf, n_sectors, out);
    run<CP_ASYNC16, 1, 8>("cp.async.cg 16 B", buf, n_sectors, out);
    run<NC, 2, 8>("ld.global.nc.u8, 2 sectors of a 64-B block", buf, n_sectors, out);
    run<NC, 4, 8>("ld.global.nc.u8, 4 sectors of a 128-B block", buf, n_sectors, out);
    run<NC, 8, 8>("ld.global.nc.u8, 8 sectors of a 256-B block", buf, n_sectors, out);
    {
        const int blocks = 148 * 8 * 4, iters = 64;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        probe_bulk16<<<blocks, 256>>>(buf, n_sectors, iters, out, 7);
        cudaEventRecord(a);
        probe_bulk16<<<blocks, 256>>>(buf, n_sectors, iters, out, 100);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double loads = (double)blocks * 256 * iters;
        printf("%-44s span=1 U=1: %6.1f G loads/s, %.3f ms, %.0f loads (%s)\n", "cp.async.bulk 16 B (TMA unit)", loads / (ms * 1e-3) / 1e9, ms,
               loads, cudaGetErrorString(cudaGetLastError()));
    }
    cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
