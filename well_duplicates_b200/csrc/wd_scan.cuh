// Device-wide exclusive prefix sum over uint32 (three short kernels:
// per-chunk sums, scan of the sums in one CTA, per-chunk rescan).  Used for
// the grid-cell offsets (K1), the ring CSR offsets (K2) and the PF rank blocks
// (K3).  out has n+1 entries; out[n] is the total.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wd {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across a CTA of SCAN_THREADS; returns
// the exclusive prefix and writes the CTA total to *total.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    __shared__ uint32_t grand;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0u;
        const uint32_t wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = wi - w;
        if (lane == 31) grand = wi;
    }
    __syncthreads();
    const uint32_t r = inc - v + warp_tot[warp];
    *total = grand;
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_chunk_sums_kernel(const uint32_t *__restrict__ in, size_t n, uint32_t *__restrict__ sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += in[base + i];
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// one CTA: sums[0..m) -> exclusive offsets in place, total to sums[m]
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_sums_kernel(uint32_t *sums, size_t m) {
    uint32_t carry = 0;
    for (size_t base = 0; base < m; base += SCAN_THREADS) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < m ? sums[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, &tot);
        if (i < m) sums[i] = ex + carry;
        carry += tot;
    }
    if (threadIdx.x == 0) sums[m] = carry;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_chunk_final_kernel(const uint32_t *__restrict__ in, size_t n, const uint32_t *__restrict__ sums,
                        size_t n_chunks, uint32_t *__restrict__ out) {
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[n_chunks];
}

// tmp must hold n_chunks+1 uint32, n_chunks = ceil(n / SCAN_CHUNK).  in and
// out may alias only if identical pointers are NOT used (out has n+1 entries).
static inline size_t scan_tmp_words(size_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 1; }

static inline cudaError_t exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n,
                                             uint32_t *tmp, cudaStream_t st, uint64_t *launches) {
    const size_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (n_chunks == 0) return cudaMemsetAsync(out, 0, sizeof(uint32_t), st);
    scan_chunk_sums_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(in, n, tmp);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(tmp, n_chunks);
    scan_chunk_final_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(in, n, tmp, n_chunks, out);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace wd
