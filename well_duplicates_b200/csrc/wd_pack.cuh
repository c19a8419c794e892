// Packed-word records in HBM (two-pass path, wd_get_seqs, exhaustive mode): per well and
// 64 symbols one 32-byte record {lo, hi, nn, meta}; meta bit 0 of the first record = PF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wd_seq.cuh"

namespace wd {

constexpr int PACK_STRIDE = 4;   // u64 per 64-symbol word group in HBM: lo, hi, nn, meta (32 B)

template <int W>
__device__ __forceinline__ void store_packed(uint64_t *dst, const PSeq<W> &s, uint64_t meta) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        ulonglong2 *q = reinterpret_cast<ulonglong2 *>(dst + (size_t)w * PACK_STRIDE);
        q[0] = make_ulonglong2(s.lo[w], s.hi[w]);
        q[1] = make_ulonglong2(s.nn[w], w == 0 ? meta : 0ull);
    }
}

template <int W>
__device__ __forceinline__ uint64_t load_packed(const uint64_t *src, PSeq<W> &s) {
    uint64_t meta = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(src + (size_t)w * PACK_STRIDE);
        const ulonglong2 a = q[0], b = q[1];
        s.lo[w] = a.x; s.hi[w] = a.y; s.nn[w] = b.x;
        if (w == 0) meta = b.y;
    }
    return meta;
}

}  // namespace wd
