// Stage 2 (K3 filter rank, K4/K5 gather-decode) and stage 3 (K6 compare +
// count, K7 publish), plus the fused gather+compare kernel that is the
// production path of wd_count.
//
// Reference semantics restated here:
//  * PF flag = filter byte & 1; offset = rank among PF wells
//    (bcl_direct_reader.py:222-253)
//  * BCL call: byte 0 -> N, else base = byte & 3 (:352-354)
//  * CBCL call: well w (or its PF rank when the block excludes non-PF wells,
//    rank -1 -> N) -> low nibble if w even else high; nibble 0 -> N, else
//    base = nibble & 3 (:303-325)
//  * a target counts only if its centre is PF; ring wells are not PF-checked;
//    N is an ordinary symbol (count_well_duplicates.py:236-262)
//  * per-tile counters as output_writer sums them (:65-106)
#include "wd_common.cuh"
#include "wd_scan.cuh"
#include "wd_seq.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace wd {

constexpr int PACK_STRIDE = 4;   // u64 per 64-symbol word group in HBM: lo, hi, nn, meta (32 B)
constexpr int MAX_ORDER = WD_MAX_SEQ_LEN;

// ============================================================================
// K3: filter bytes -> PF bit mask + block ranks
// ============================================================================
// one thread per 64 wells; the filter buffer is zero-padded to a multiple of 64
__global__ void __launch_bounds__(256)
filter_mask_kernel(const uint8_t *__restrict__ filt, uint32_t n_blocks, uint64_t *__restrict__ mask,
                   uint32_t *__restrict__ cnt) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(filt + (size_t)b * 64);
    uint64_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldg(src + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bit0 of each of the 4 bytes -> 4 adjacent bits
            const uint32_t bits = (((w[k] & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
            m |= (uint64_t)bits << (q * 16 + k * 4);
        }
    }
    mask[b] = m;
    cnt[b] = (uint32_t)__popcll(m);
}

__device__ __forceinline__ int pf_rank(const TileDesc &d, uint32_t well) {
    const uint64_t m = __ldg(d.pfmask + (well >> 6));
    const int b = well & 63;
    if (!((m >> b) & 1ull)) return -1;
    return (int)(__ldg(d.pfrank + (well >> 6)) + (uint32_t)__popcll(m & ((1ull << b) - 1ull)));
}

// the reference's filter_offsets list (parity hook)
__global__ void __launch_bounds__(256)
filter_expand_kernel(TileDesc d, int32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.n) out[i] = pf_rank(d, i);
}

int filter_rank(wd_ctx *ctx, const int *slot_ids, int n) {
    cudaStream_t st = ctx->stream;
    for (int k = 0; k < n; ++k) {
        TileSlot &s = ctx->slots[slot_ids[k]];
        if (s.rank_valid) continue;
        if (!s.filter_set) WD_FAIL(WD_E_ARG, "tile slot %d has no filter loaded", slot_ids[k]);
        const uint32_t nb = (s.n + 63) / 64;
        if (s.mapped_filter) {
            // K3 reads whole zero-padded 64-byte blocks: bring a mapped filter into HBM first
            const size_t fstride = ((size_t)s.n + 255) & ~(size_t)255;
            WD_CUDA(cudaMemcpyAsync(s.filter.p, s.mapped_filter_host, s.n, cudaMemcpyHostToDevice, st));
            if (fstride > s.n) WD_CUDA(cudaMemsetAsync(s.filter.as<uint8_t>() + s.n, 0, fstride - s.n, st));
        }
        WD_TRY(s.pfmask.reserve((size_t)nb * 8));
        WD_TRY(s.pfrank.reserve(((size_t)nb + 1) * 4));
        WD_TRY(s.pfcount_dev.reserve((size_t)nb * 4));
        WD_TRY(ctx->scan_tmp.reserve(scan_tmp_words(nb) * 4));
        filter_mask_kernel<<<(nb + 255) / 256, 256, 0, st>>>(s.filter.as<uint8_t>(), nb, s.pfmask.as<uint64_t>(),
                                                              s.pfcount_dev.as<uint32_t>());
        ctx->launches++;
        WD_CUDA(exclusive_scan_u32(s.pfcount_dev.as<uint32_t>(), s.pfrank.as<uint32_t>(), nb,
                                   ctx->scan_tmp.as<uint32_t>(), st, &ctx->launches));
        s.rank_valid = true;
    }
    return WD_OK;
}

static TileDesc make_desc(const TileSlot &s) {
    TileDesc d;
    d.planes = s.mapped ? s.mapped : s.planes.as<uint8_t>();
    d.stride = s.stride;
    d.filter = s.mapped_filter ? s.mapped_filter : s.filter.as<uint8_t>();
    d.pfmask = s.pfmask.as<uint64_t>();
    d.pfrank = s.pfrank.as<uint32_t>();
    d.kind = s.kind_dev.as<uint8_t>();
    d.n = s.n;
    d.flags = s.rank_valid ? 1u : 0u;
    d.head_stride = ((unsigned long long)s.n + 255ull) & ~255ull;
    d.head_delta = (unsigned long long)(uintptr_t)s.head.as<uint8_t>() - (unsigned long long)(uintptr_t)d.planes;
    return d;
}

int filter_offsets(wd_ctx *ctx, int slot, int32_t *offsets, uint32_t *passing) {
    TileSlot &s = ctx->slots[slot];
    WD_TRY(filter_rank(ctx, &slot, 1));
    cudaStream_t st = ctx->stream;
    WD_TRY(ctx->gs_codes.reserve((size_t)s.n * 4));
    filter_expand_kernel<<<(s.n + 255) / 256, 256, 0, st>>>(make_desc(s), ctx->gs_codes.as<int32_t>());
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    WD_CUDA(cudaMemcpyAsync(offsets, ctx->gs_codes.p, (size_t)s.n * 4, cudaMemcpyDeviceToHost, st));
    uint32_t pf = 0;
    WD_CUDA(cudaMemcpyAsync(&pf, s.pfrank.as<uint32_t>() + (s.n + 63) / 64, 4, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    if (passing) *passing = pf;
    return WD_OK;
}

// ============================================================================
// K4 / K5: gather-decode one well into packed words
// ============================================================================
// s_off[p]  = byte offset of the plane that supplies sequence position p
// s_kind[p] = WD_PLANE_* of that plane (same for every tile of a launch)
// one base call -> raw code: 0 = no-call, otherwise base = code & 3
template <bool ALL_BCL>
__device__ __forceinline__ uint32_t load_call(const TileDesc &d, uint32_t well, int rank, unsigned long long off,
                                              int kind) {
    if (ALL_BCL || kind == WD_PLANE_BCL) return __ldg(d.planes + off + well);
    const int wi = kind == WD_PLANE_CBCL_EXCL ? rank : (int)well;
    if (wi < 0) return 0u;                        // not PF: the block has no entry for it -> N
    const uint32_t byte = __ldg(d.planes + off + ((uint32_t)wi >> 1));
    return (wi & 1) ? (byte >> 4) : (byte & 15u);
}

// N (8 or 16) consecutive sequence positions p .. p+N-1 (those >= len
// contribute nothing) -> N-bit groups of the three planes.  All N loads are
// issued before any is consumed (memory-level parallelism).
template <bool ALL_BCL, int N>
__device__ __forceinline__ void decode_n(const TileDesc &d, uint32_t well, int rank, const unsigned long long *s_off,
                                         const uint8_t *s_kind, int p, int len, uint32_t &glo, uint32_t &ghi,
                                         uint32_t &gnn) {
    uint32_t code[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        code[j] = 4u;                            // beyond the sequence: no bit in any plane
        if (p + j < len) code[j] = load_call<ALL_BCL>(d, well, rank, s_off[p + j], ALL_BCL ? 0 : s_kind[p + j]);
    }
    glo = ghi = gnn = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t b = code[j];
        glo |= (b & 1u) << j;
        ghi |= ((b >> 1) & 1u) << j;
        gnn |= (b == 0u ? 1u : 0u) << j;
    }
}

template <int W>
__device__ __forceinline__ void pseq_or_group(PSeq<W> &q, int p, uint32_t glo, uint32_t ghi, uint32_t gnn) {
    const int w = p >> 6, sh = p & 63;           // groups start at multiples of their size: none straddles a word
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == w) {
            q.lo[i] |= (uint64_t)glo << sh;
            q.hi[i] |= (uint64_t)ghi << sh;
            q.nn[i] |= (uint64_t)gnn << sh;
        }
    }
}

template <int W, bool ALL_BCL>
__device__ __forceinline__ void decode_well(const TileDesc &d, uint32_t well, const unsigned long long *s_off,
                                            const uint8_t *s_kind, int len, PSeq<W> &out) {
    int rank = 0;
    if (!ALL_BCL && (d.flags & 1u)) rank = pf_rank(d, well);
    pseq_clear(out);
    for (int p = 0; p < len; p += 8) {
        uint32_t glo, ghi, gnn;
        decode_n<ALL_BCL, 8>(d, well, rank, s_off, s_kind, p, len, glo, ghi, gnn);
        pseq_or_group<W>(out, p, glo, ghi, gnn);
    }
}

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

__device__ __forceinline__ void load_order(const unsigned long long *g_off, const uint8_t *g_kind, int len,
                                           unsigned long long *s_off, uint8_t *s_kind) {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        s_off[i] = g_off[i];
        s_kind[i] = g_kind[i];
    }
    __syncthreads();
}

// one thread per (tile, slot): packed[tile][slot][W][4]
template <int W, bool ALL_BCL>
__global__ void __launch_bounds__(256)
gather_pack_kernel(const TileDesc *__restrict__ descs, const uint32_t *__restrict__ slot_well, uint32_t n_slots,
                   const unsigned long long *__restrict__ g_off, const uint8_t *__restrict__ g_kind, int len,
                   uint64_t *__restrict__ packed) {
    __shared__ unsigned long long s_off[MAX_ORDER];
    __shared__ uint8_t s_kind[MAX_ORDER];
    load_order(g_off, g_kind, len, s_off, s_kind);
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const TileDesc d = descs[blockIdx.y];
    const uint32_t well = slot_well ? __ldg(slot_well + s) : s;     // null: every well in index order (exhaustive mode)
    PSeq<W> q;
    decode_well<W, ALL_BCL>(d, well, s_off, s_kind, len, q);
    const uint64_t meta = __ldg(d.filter + well) & 1u;
    store_packed<W>(packed + ((size_t)blockIdx.y * n_slots + s) * (size_t)(W * PACK_STRIDE), q, meta);
}

// packed -> one byte per symbol (0..3 ACGT, 4 N) for wd_get_seqs
template <int W>
__global__ void __launch_bounds__(256)
unpack_codes_kernel(const uint64_t *__restrict__ packed, uint32_t n_idx, int len, uint8_t *__restrict__ codes,
                    uint8_t *__restrict__ pf) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_idx) return;
    PSeq<W> q;
    const uint64_t meta = load_packed<W>(packed + (size_t)s * (W * PACK_STRIDE), q);
    pf[s] = (uint8_t)(meta & 1u);
    for (int p = 0; p < len; ++p) {
        const unsigned c = pseq_get<W>(q, p);
        codes[(size_t)s * len + p] = (uint8_t)((c & 4u) ? 4u : c);
    }
}

// ============================================================================
// K6: compare + count
// ============================================================================
struct CountArgs {
    const TileDesc *descs;
    const uint32_t *tgt_off, *slot_well, *slot_csr, *level_len, *visit;
    const uint8_t *slot_level;
    const unsigned long long *g_off;
    const uint8_t *g_kind;
    const uint64_t *packed;          // two-pass only
    int32_t *per_target;             // may be null
    unsigned long long *counters;    // [tiles][1+5L]
    int32_t *dup_rows;               // may be null: (tile, target, csr position, distance)
    unsigned long long *dup_count;
    unsigned long long dup_cap;
    uint32_t t, n_slots;
    int levels, len, e, hamming;
    int step0, step1;                // fused kernel: cycles read per round (first, later), 1..16
    int cchunk;                      // fused kernel: centre cycles decoded per warp-wide load (8, 16 or 32)
    int n_head;                      // fused kernel: positions 0..n_head-1 are read from the tile's head planes in HBM
};

// Per-warp tallies -> per_target row and the CTA's shared counters.
template <int LMAX>
__device__ __forceinline__ void finish_target(const CountArgs &a, uint32_t tile, uint32_t t, int lane, bool valid,
                                              const uint32_t *dups, uint32_t *s_cnt) {
    const int L = a.levels;
    const int row = 1 + 2 * L;
    if (a.per_target != nullptr) {
        int32_t *pt = a.per_target + ((size_t)tile * a.t + t) * row;
        if (lane == 0) pt[0] = valid ? 1 : 0;
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
            if (l < L && lane == l) {
                pt[1 + 2 * l] = valid ? (int32_t)dups[l] : 0;
                pt[2 + 2 * l] = valid ? (int32_t)__ldg(a.level_len + (size_t)t * L + l) : 0;
            }
        }
    }
    if (!valid) return;
    // AccO: a hit at this level or further in; AccI: at this level or further out
    // (count_well_duplicates.py:77-89)
    uint32_t hit_mask = 0;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
        if (l < L && dups[l]) hit_mask |= 1u << l;
    if (lane == 0) atomicAdd(&s_cnt[0], 1u);
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
        if (l < L && lane == l) {
            uint32_t *c = s_cnt + 1 + 5 * l;
            atomicAdd(c + 0, __ldg(a.level_len + (size_t)t * L + l));
            if (dups[l]) {
                atomicAdd(c + 1, dups[l]);
                atomicAdd(c + 2, 1u);
            }
            if (hit_mask & ((2u << l) - 1u)) atomicAdd(c + 3, 1u);
            if (hit_mask >> l) atomicAdd(c + 4, 1u);
        }
    }
}

template <int W>
__device__ __forceinline__ void log_dup(const CountArgs &a, uint32_t tile, uint32_t t, uint32_t slot,
                                        const PSeq<W> &c, const PSeq<W> &b) {
    const unsigned long long pos = atomicAdd(a.dup_count, 1ull);
    if (pos < a.dup_cap) {
        int32_t *r = a.dup_rows + pos * 4;
        r[0] = (int32_t)tile;
        r[1] = (int32_t)t;
        r[2] = (int32_t)__ldg(a.slot_csr + slot);
        r[3] = exact_distance<W>(c, b, a.len, a.hamming != 0);
    }
}

__device__ __forceinline__ void flush_counters(uint32_t *s_cnt, unsigned long long *dst, int n) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(dst + i, (unsigned long long)s_cnt[i]);
}

constexpr int CNT_WARPS = 8;

// two-pass flavour: reads the packed words K4/K5 left in HBM.  One warp per
// (tile, target); lanes stride over the target's ring slots.
template <int W, int LMAX>
__global__ void __launch_bounds__(CNT_WARPS * 32)
compare_count_kernel(CountArgs a) {
    __shared__ uint32_t s_cnt[1 + 5 * LMAX];
    for (int i = threadIdx.x; i < 1 + 5 * LMAX; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.y;
    const uint32_t t = blockIdx.x * CNT_WARPS + (threadIdx.x >> 5);
    if (t < a.t) {
        const uint32_t s0 = __ldg(a.tgt_off + t), s1 = __ldg(a.tgt_off + t + 1);
        const uint64_t *tp = a.packed + (size_t)tile * a.n_slots * (W * PACK_STRIDE);
        PSeq<W> c;
        const uint64_t meta = load_packed<W>(tp + (size_t)s0 * (W * PACK_STRIDE), c);
        const bool valid = (meta & 1ull) != 0;
        uint32_t dups[LMAX];
#pragma unroll
        for (int l = 0; l < LMAX; ++l) dups[l] = 0;
        if (valid) {
            for (uint32_t base = s0 + 1; base < s1; base += 32) {
                const uint32_t s = base + lane;
                bool dup = false;
                int lvl = 0;
                if (s < s1) {
                    PSeq<W> b;
                    load_packed<W>(tp + (size_t)s * (W * PACK_STRIDE), b);
                    lvl = __ldg(a.slot_level + s);
                    dup = is_duplicate<W>(c, b, a.len, a.e, a.hamming != 0);
                    if (dup && a.dup_rows != nullptr) log_dup<W>(a, tile, t, s, c, b);
                }
#pragma unroll
                for (int l = 0; l < LMAX; ++l)
                    if (l < a.levels) dups[l] += __popc(__ballot_sync(0xffffffffu, dup && lvl == l + 1));
            }
        }
        finish_target<LMAX>(a, tile, t, lane, valid, dups, s_cnt);
    }
    flush_counters(s_cnt, a.counters + (size_t)tile * (1 + 5 * a.levels), 1 + 5 * a.levels);
}

// fused flavour (production): the warp gathers and decodes its target's wells
// straight from the planes, compares in registers and never writes the packed
// words.  What it reads is decided symbol by symbol:
//  * targets whose centre fails the filter are skipped before any plane byte
//    is read;
//  * ring wells are read a few cycles at a time, one well per lane, and fed to
//    the incremental edit-distance programme of wd_seq.cuh (PrefixDP; a running
//    mismatch count for --hamming / e < 2).  A well stops being read as soon as
//    its prefix proves dist > e -- 96 % of unrelated reads after 6 symbols --
//    so the later planes are touched only around real duplicates;
//  * the centre is decoded by the whole warp (lane = cycle, three ballots turn
//    the calls into bit-plane words), 8-32 cycles at a time and only as
//    far ahead as the programme needs (k = e/2 symbols past the ring wells): a
//    target without duplicates never reads its centre beyond the first chunks.
constexpr int FUSED_TPB = 64;    // targets per CTA; its 8 warps pull them from a shared counter

// raw call (0 = no-call, else base = raw & 3) -> symbol 0..3, 4 = N
__device__ __forceinline__ uint32_t call_symbol(uint32_t raw) { return raw == 0u ? 4u : (raw & 3u); }

// n <= 16 bits of a W-word bit string starting at bit p
template <int W>
__device__ __forceinline__ uint32_t bits_at(const uint64_t *plane, int p, int n) {
    const int w = p >> 6, sh = p & 63;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == w) v |= plane[i] >> sh;
        if (i == w + 1 && sh != 0) v |= plane[i] << (64 - sh);
    }
    return (uint32_t)v & ((1u << n) - 1u);
}

// NMAX calls of one well at sequence positions p .. p+n-1, all loads in flight together
template <bool ALL_BCL, int NMAX>
__device__ __forceinline__ void load_calls(const TileDesc &d, uint32_t well, int rank, const unsigned long long *s_off,
                                           const uint8_t *s_kind, int p, int n, uint32_t (&raw)[NMAX]) {
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        raw[j] = 0u;
        if (j < n) raw[j] = load_call<ALL_BCL>(d, well, rank, s_off[p + j], ALL_BCL ? 0 : s_kind[p + j]);
    }
}

template <int W, bool ALL_BCL, int NMAX>
__device__ __forceinline__ bool ring_round(const TileDesc &d, uint32_t well, int rank, const unsigned long long *s_off,
                                           const uint8_t *s_kind, const PSeq<W> &c, int known_c, int len, int p, int n,
                                           int k, int e, bool ham_like, PrefixDP<W> &dp, int &mism) {
    uint32_t raw[NMAX];
    load_calls<ALL_BCL, NMAX>(d, well, rank, s_off, s_kind, p, n, raw);
    if (ham_like) {
        uint32_t glo = 0, ghi = 0, gnn = 0;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < n) {
                const uint32_t sym = call_symbol(raw[j]);
                glo |= (sym & 1u) << j;
                ghi |= ((sym >> 1) & 1u) << j;
                gnn |= (sym >> 2) << j;
            }
        }
        mism += __popc((bits_at<W>(c.lo, p, n) ^ glo) | (bits_at<W>(c.hi, p, n) ^ ghi) | (bits_at<W>(c.nn, p, n) ^ gnn));
        return mism <= e;
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
        if (j < n) pdp_step<W>(dp, c, known_c, len, p + j, k, call_symbol(raw[j]));
    return pdp_band_min<W>(dp, len, p + n, k) <= e;
}

template <int W, int LMAX, bool ALL_BCL>
__global__ void __launch_bounds__(CNT_WARPS * 32)
fused_count_kernel(CountArgs a) {
    __shared__ unsigned long long s_off[MAX_ORDER];
    __shared__ uint8_t s_kind[MAX_ORDER];
    __shared__ uint32_t s_cnt[1 + 5 * LMAX];
    __shared__ uint32_t s_next;
    for (int i = threadIdx.x; i < 1 + 5 * LMAX; i += blockDim.x) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    load_order(a.g_off, a.g_kind, a.len, s_off, s_kind);
    const int lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.y;
    const TileDesc d = a.descs[tile];
    const int len = a.len, e = a.e;
    if (a.n_head) {
        // this CTA serves one tile: point its first positions at the tile's head planes
        for (int i = threadIdx.x; i < a.n_head; i += blockDim.x) s_off[i] = d.head_delta + (unsigned long long)i * d.head_stride;
        __syncthreads();
    }
    // Levenshtein <= 1 <=> Hamming <= 1 on equal lengths (an indel pair costs 2)
    const bool ham_like = a.hamming != 0 || e < 2;
    const int k = ham_like ? 0 : (e >> 1);
    const bool read_nothing = e < 0 || e >= len;       // no pair / every pair is a duplicate
    const uint32_t t_begin = blockIdx.x * FUSED_TPB;
    const uint32_t t_end = min(t_begin + FUSED_TPB, a.t);
    // Targets differ a lot in cost (a failed centre costs one byte, a real
    // duplicate keeps its warp reading to the last cycle), so warps take the
    // next target when they are done instead of owning a fixed one.
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = t_begin + atomicAdd(&s_next, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= t_end) break;
        if (a.visit) t = __ldg(a.visit + t);              // spatial visiting order; results go by target ordinal
        const uint32_t s0 = __ldg(a.tgt_off + t), s1 = __ldg(a.tgt_off + t + 1);
        const uint32_t centre = __ldg(a.slot_well + s0);
        const bool valid = (__ldg(d.filter + centre) & 1u) != 0;
        uint32_t dups[LMAX];
#pragma unroll
        for (int l = 0; l < LMAX; ++l) dups[l] = 0;
        if (valid) {
            PSeq<W> c;
            pseq_clear(c);
            int known_c = 0;
            int crank = 0;
            if (!ALL_BCL && (d.flags & 1u)) crank = pf_rank(d, centre);
            for (uint32_t base = s0 + 1; base < s1; base += 32) {
                const uint32_t s = base + lane;
                const bool mine = s < s1;
                uint32_t well = 0;
                int lvl = 0, rank = 0;
                if (mine) {
                    well = __ldg(a.slot_well + s);
                    lvl = __ldg(a.slot_level + s);
                    if (!ALL_BCL && (d.flags & 1u)) rank = pf_rank(d, well);
                }
                PrefixDP<W> dp;
                pdp_init(dp);
                int mism = 0;
                bool alive = mine && e >= 0;
                int p = read_nothing ? len : 0;
                while (p < len) {
                    if (!__any_sync(0xffffffffu, alive)) break;
                    const int n = min(p == 0 ? a.step0 : a.step1, len - p);
                    // ---- centre: lane = cycle, as far as this round looks ahead ----------
                    const int need = min(len, p + n + k);
                    while (known_c < need) {
                        const int q = known_c + lane;
                        uint32_t sym = 0u;
                        if (lane < a.cchunk && q < len)
                            sym = call_symbol(load_call<ALL_BCL>(d, centre, crank, s_off[q], ALL_BCL ? 0 : s_kind[q]));
                        const uint32_t glo = __ballot_sync(0xffffffffu, sym & 1u);
                        const uint32_t ghi = __ballot_sync(0xffffffffu, sym & 2u);
                        const uint32_t gnn = __ballot_sync(0xffffffffu, sym & 4u);
                        pseq_or_group<W>(c, known_c, glo, ghi, gnn);     // chunks never straddle a word
                        known_c = min(len, known_c + a.cchunk);
                    }
                    // ---- ring wells: lane = well ---------------------------------------------
                    if (alive) {
                        if (n > 8) alive = ring_round<W, ALL_BCL, 16>(d, well, rank, s_off, s_kind, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                        else if (n > 4) alive = ring_round<W, ALL_BCL, 8>(d, well, rank, s_off, s_kind, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                        else if (n > 2) alive = ring_round<W, ALL_BCL, 4>(d, well, rank, s_off, s_kind, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                        else alive = ring_round<W, ALL_BCL, 2>(d, well, rank, s_off, s_kind, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                    }
                    p += n;
                }
                const bool dup = alive;            // survived to p == len: dist <= e
#pragma unroll
                for (int l = 0; l < LMAX; ++l)
                    if (l < a.levels) dups[l] += __popc(__ballot_sync(0xffffffffu, dup && lvl == l + 1));
            }
        }
        finish_target<LMAX>(a, tile, t, lane, valid, dups, s_cnt);
    }
    flush_counters(s_cnt, a.counters + (size_t)tile * (1 + 5 * a.levels), 1 + 5 * a.levels);
}

// ============================================================================
// K7: place tile counters into the all-reduce buffer
// ============================================================================
__global__ void publish_kernel(const unsigned long long *__restrict__ counters, const int32_t *__restrict__ tile_row,
                               const int32_t *__restrict__ lane_row, int n_tiles, int width,
                               unsigned long long *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * width) return;
    const int tile = i / width, k = i % width;
    const unsigned long long v = counters[i];
    if (tile_row[tile] >= 0) out[(size_t)tile_row[tile] * width + k] = v;
    if (lane_row[tile] >= 0 && v) atomicAdd(out + (size_t)lane_row[tile] * width + k, v);
}

// ============================================================================
// host side
// ============================================================================
static int words_for(int len) {
    if (len <= 64) return 1;
    if (len <= 128) return 2;
    if (len <= 256) return 4;
    if (len <= 512) return 8;
    return 16;
}

// plane order -> per-position byte offsets and kinds; checks that every tile
// of the batch agrees on stride / kinds (they share one launch).
static int prepare_order(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int len, bool *all_bcl,
                         bool *any_excl) {
    if (len < 1 || len > WD_MAX_SEQ_LEN)
        WD_FAIL(WD_E_ARG, "compared sequence length %d is outside 1..%d", len, WD_MAX_SEQ_LEN);
    if (first_slot < 0 || n_tiles < 1 || (size_t)first_slot + n_tiles > ctx->slots.size())
        WD_FAIL(WD_E_ARG, "tile slots %d..%d have not been begun", first_slot, first_slot + n_tiles - 1);
    const TileSlot &s0 = ctx->slots[first_slot];
    std::vector<unsigned long long> off(len);
    std::vector<uint8_t> kind(len);
    *all_bcl = true;
    *any_excl = false;
    for (int p = 0; p < len; ++p) {
        const int pl = order[p];
        if (pl < 0 || pl >= s0.n_planes) WD_FAIL(WD_E_ARG, "plane %d out of range (slot has %d planes)", pl, s0.n_planes);
        off[p] = (unsigned long long)pl * s0.stride;
        kind[p] = s0.kind[pl];
        if (kind[p] == WD_PLANE_EMPTY) WD_FAIL(WD_E_ARG, "plane %d of tile slot %d was never loaded", pl, first_slot);
        if (kind[p] != WD_PLANE_BCL) *all_bcl = false;
        if (kind[p] == WD_PLANE_CBCL_EXCL) *any_excl = true;
    }
    for (int k = 0; k < n_tiles; ++k) {
        const TileSlot &s = ctx->slots[first_slot + k];
        if (s.n == 0 || !s.filter_set) WD_FAIL(WD_E_ARG, "tile slot %d is not fully loaded", first_slot + k);
        if (s.stride != s0.stride || s.n_planes != s0.n_planes)
            WD_FAIL(WD_E_ARG, "tile slots of one batch must have the same plane count and stride");
        for (int p = 0; p < len; ++p)
            if (s.kind[order[p]] != kind[p])
                WD_FAIL(WD_E_ARG, "tile slot %d: plane %d has a different format than in slot %d", first_slot + k,
                        order[p], first_slot);
    }
    cudaStream_t st = ctx->stream;
    WD_TRY(ctx->order_dev.reserve((size_t)MAX_ORDER * 9));
    WD_CUDA(cudaMemcpyAsync(ctx->order_dev.p, off.data(), (size_t)len * 8, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8, kind.data(), (size_t)len,
                            cudaMemcpyHostToDevice, st));
    return WD_OK;
}

int upload_descs(wd_ctx *ctx, int first_slot, int n_tiles) {
    std::vector<TileDesc> h(n_tiles);
    cudaStream_t st = ctx->stream;
    for (int k = 0; k < n_tiles; ++k) {
        TileSlot &s = ctx->slots[first_slot + k];
        if (s.kind_dirty) {
            WD_TRY(s.kind_dev.reserve((size_t)s.n_planes));
            WD_CUDA(cudaMemcpyAsync(s.kind_dev.p, s.kind.data(), (size_t)s.n_planes, cudaMemcpyHostToDevice, st));
            s.kind_dirty = false;
        }
        h[k] = make_desc(s);
    }
    WD_TRY(ctx->descs.reserve((size_t)n_tiles * sizeof(TileDesc)));
    WD_CUDA(cudaMemcpyAsync(ctx->descs.p, h.data(), (size_t)n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    return WD_OK;
}

template <int W, bool ALL_BCL>
static void launch_gather(wd_ctx *ctx, const TileDesc *descs, const uint32_t *slot_well, uint32_t n_slots, int n_tiles,
                          int len, uint64_t *packed) {
    const unsigned long long *g_off = ctx->order_dev.as<unsigned long long>();
    const uint8_t *g_kind = ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8;
    dim3 grid((n_slots + 255) / 256, n_tiles);
    gather_pack_kernel<W, ALL_BCL><<<grid, 256, 0, ctx->stream>>>(descs, slot_well, n_slots, g_off, g_kind, len, packed);
    ctx->launches++;
}

template <int W>
static void launch_gather_w(wd_ctx *ctx, bool all_bcl, const TileDesc *descs, const uint32_t *slot_well,
                            uint32_t n_slots, int n_tiles, int len, uint64_t *packed) {
    if (all_bcl) launch_gather<W, true>(ctx, descs, slot_well, n_slots, n_tiles, len, packed);
    else launch_gather<W, false>(ctx, descs, slot_well, n_slots, n_tiles, len, packed);
}

static void launch_gather_any(wd_ctx *ctx, int words, bool all_bcl, const TileDesc *descs, const uint32_t *slot_well,
                              uint32_t n_slots, int n_tiles, int len, uint64_t *packed) {
    switch (words) {
        case 1: launch_gather_w<1>(ctx, all_bcl, descs, slot_well, n_slots, n_tiles, len, packed); break;
        case 2: launch_gather_w<2>(ctx, all_bcl, descs, slot_well, n_slots, n_tiles, len, packed); break;
        case 4: launch_gather_w<4>(ctx, all_bcl, descs, slot_well, n_slots, n_tiles, len, packed); break;
        case 8: launch_gather_w<8>(ctx, all_bcl, descs, slot_well, n_slots, n_tiles, len, packed); break;
        default: launch_gather_w<16>(ctx, all_bcl, descs, slot_well, n_slots, n_tiles, len, packed); break;
    }
}

int get_seqs(wd_ctx *ctx, int slot, const int64_t *indices, uint32_t n_idx, const int32_t *order, int seq_len,
             uint8_t *codes, uint8_t *pf) {
    if (slot < 0 || (size_t)slot >= ctx->slots.size() || ctx->slots[slot].n == 0)
        WD_FAIL(WD_E_ARG, "tile slot %d has not been begun", slot);
    TileSlot &s = ctx->slots[slot];
    if (n_idx == 0) return WD_OK;
    // bcl_direct_reader.py:186-192
    int64_t mx = indices[0], mn = indices[0];
    for (uint32_t i = 1; i < n_idx; ++i) {
        mx = indices[i] > mx ? indices[i] : mx;
        mn = indices[i] < mn ? indices[i] : mn;
    }
    if (mx >= (int64_t)s.n)
        WD_FAIL(WD_E_INDEX, "Requested cluster %lld is out of range.  Highest on this tile is %u.", (long long)mx, s.n - 1);
    if (mn < 0) WD_FAIL(WD_E_INDEX, "Requested cluster %lld is a negative number.", (long long)mn);
    cudaStream_t st = ctx->stream;
    std::vector<uint32_t> wells(n_idx);
    for (uint32_t i = 0; i < n_idx; ++i) wells[i] = (uint32_t)indices[i];
    WD_TRY(ctx->gs_idx.reserve((size_t)n_idx * 4));
    WD_CUDA(cudaMemcpyAsync(ctx->gs_idx.p, wells.data(), (size_t)n_idx * 4, cudaMemcpyHostToDevice, st));
    if (!s.filter_set) WD_FAIL(WD_E_ARG, "tile slot %d has no filter loaded", slot);
    if (seq_len == 0) {
        // zero-length range: only the flags
        WD_TRY(filter_rank(ctx, &slot, 1));
        std::vector<uint8_t> f(s.n);
        if (s.mapped_filter_host) memcpy(f.data(), s.mapped_filter_host, s.n);
        else WD_CUDA(cudaMemcpyAsync(f.data(), s.filter.p, s.n, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < n_idx; ++i) pf[i] = f[wells[i]] & 1;
        return WD_OK;
    }
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, slot, 1, order, seq_len, &all_bcl, &any_excl));
    if (any_excl) WD_TRY(filter_rank(ctx, &slot, 1));
    WD_TRY(upload_descs(ctx, slot, 1));
    const int words = words_for(seq_len);
    WD_TRY(ctx->gs_packed.reserve((size_t)n_idx * words * PACK_STRIDE * 8));
    WD_TRY(ctx->gs_codes.reserve((size_t)n_idx * seq_len + n_idx));
    launch_gather_any(ctx, words, all_bcl, ctx->descs.as<TileDesc>(), ctx->gs_idx.as<uint32_t>(), n_idx, 1, seq_len,
                      ctx->gs_packed.as<uint64_t>());
    uint8_t *d_codes = ctx->gs_codes.as<uint8_t>();
    uint8_t *d_pf = d_codes + (size_t)n_idx * seq_len;
    const unsigned blocks = (n_idx + 255) / 256;
    switch (words) {
        case 1: unpack_codes_kernel<1><<<blocks, 256, 0, st>>>(ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 2: unpack_codes_kernel<2><<<blocks, 256, 0, st>>>(ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 4: unpack_codes_kernel<4><<<blocks, 256, 0, st>>>(ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 8: unpack_codes_kernel<8><<<blocks, 256, 0, st>>>(ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        default: unpack_codes_kernel<16><<<blocks, 256, 0, st>>>(ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
    }
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    WD_CUDA(cudaMemcpyAsync(codes, d_codes, (size_t)n_idx * seq_len, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaMemcpyAsync(pf, d_pf, (size_t)n_idx, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    return WD_OK;
}

template <int W, int LMAX>
static void launch_count(wd_ctx *ctx, const CountArgs &a, int n_tiles, int mode, bool all_bcl) {
    dim3 grid((a.t + CNT_WARPS - 1) / CNT_WARPS, n_tiles);
    dim3 fgrid((a.t + FUSED_TPB - 1) / FUSED_TPB, n_tiles);
    if (mode == 1) {
        compare_count_kernel<W, LMAX><<<grid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
    } else if (all_bcl) {
        fused_count_kernel<W, LMAX, true><<<fgrid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
    } else {
        fused_count_kernel<W, LMAX, false><<<fgrid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
}

template <int W>
static void launch_count_w(wd_ctx *ctx, const CountArgs &a, int n_tiles, int mode, bool all_bcl) {
    if (a.levels <= 5) launch_count<W, 5>(ctx, a, n_tiles, mode, all_bcl);
    else launch_count<W, WD_MAX_LEVELS>(ctx, a, n_tiles, mode, all_bcl);
}

int count_async(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e, int hamming,
                int mode, int want_per_target) {
    TargetList &tl = ctx->targets;
    if (tl.t == 0) WD_FAIL(WD_E_ARG, "wd_count: no target list loaded");
    if (mode != 0 && mode != 1) WD_FAIL(WD_E_ARG, "wd_count: mode must be 0 (fused) or 1 (two-pass)");
    if (n_tiles > 65535) WD_FAIL(WD_E_ARG, "wd_count: at most 65535 tiles per call");
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, first_slot, n_tiles, order, seq_len, &all_bcl, &any_excl));
    for (int k = 0; k < n_tiles; ++k) {
        const TileSlot &s = ctx->slots[first_slot + k];
        if (tl.max_well >= s.n)
            WD_FAIL(WD_E_INDEX, "Requested cluster %u is out of range.  Highest on this tile is %u.", tl.max_well, s.n - 1);
    }
    if (any_excl) {
        for (int k = 0; k < n_tiles; ++k) {
            const int id = first_slot + k;
            WD_TRY(filter_rank(ctx, &id, 1));
        }
    }
    WD_TRY(upload_descs(ctx, first_slot, n_tiles));
    cudaStream_t st = ctx->stream;
    const int L = tl.levels;
    const size_t width = 1 + 5 * (size_t)L;
    WD_TRY(ctx->counters.reserve((size_t)n_tiles * width * 8));
    WD_CUDA(cudaMemsetAsync(ctx->counters.p, 0, (size_t)n_tiles * width * 8, st));
    if (want_per_target) WD_TRY(ctx->per_target.reserve((size_t)n_tiles * tl.t * (1 + 2 * L) * 4));
    const int words = words_for(seq_len);

    CountArgs a;
    a.descs = ctx->descs.as<TileDesc>();
    a.tgt_off = tl.tgt_off.as<uint32_t>();
    a.slot_well = tl.slot_well.as<uint32_t>();
    a.slot_csr = tl.slot_csr.as<uint32_t>();
    a.level_len = tl.level_len.as<uint32_t>();
    a.visit = getenv("WELLDUP_NO_VISIT_ORDER") ? nullptr : tl.visit.as<uint32_t>();
    a.slot_level = tl.slot_level.as<uint8_t>();
    a.g_off = ctx->order_dev.as<unsigned long long>();
    a.g_kind = ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8;
    a.packed = nullptr;
    a.per_target = want_per_target ? ctx->per_target.as<int32_t>() : nullptr;
    a.counters = ctx->counters.as<unsigned long long>();
    a.dup_rows = nullptr;
    a.dup_count = nullptr;
    a.dup_cap = 0;
    a.t = tl.t;
    a.n_slots = tl.n_slots;
    a.levels = L;
    a.len = seq_len;
    a.e = e;
    a.hamming = hamming;
    // early-exit schedule of the fused kernel (cycles read per round: first, later), from the sweep
    // in profiles/r01_early_exit_sweep.txt: short rounds win -- the traffic saved by dropping a well
    // sooner outweighs the extra dependent round trips
    // early-exit schedule (cycles read per round: first, later), from the sweeps in profiles/: planes in
    // HBM are bound by instruction issue and latency -> few long rounds; planes pulled across PCIe
    // (wd_tile_map_host) are bound by the number of sector requests -> many short rounds
    const bool over_pcie = ctx->slots[first_slot].mapped != nullptr;
    a.step0 = over_pcie ? 4 : 8;
    a.step1 = over_pcie ? 1 : 2;
    a.cchunk = over_pcie ? 8 : 16;
    if (const char *cc = getenv("WELLDUP_CENTRE_CHUNK")) {
        const int v = atoi(cc);
        if (v == 8 || v == 16 || v == 32) a.cchunk = v;
    }
    if (const char *sch = getenv("WELLDUP_STEPS")) {
        int s0 = 0, s1 = 0;
        if (sscanf(sch, "%d,%d", &s0, &s1) == 2 && s0 >= 1 && s0 <= 16 && s1 >= 1 && s1 <= 16) {
            a.step0 = s0;
            a.step1 = s1;
        }
    }

    if (mode == 1) {
        WD_TRY(ctx->packed.reserve((size_t)n_tiles * tl.n_slots * words * PACK_STRIDE * 8));
        launch_gather_any(ctx, words, all_bcl, a.descs, a.slot_well, tl.n_slots, n_tiles, seq_len,
                          ctx->packed.as<uint64_t>());
        a.packed = ctx->packed.as<uint64_t>();
        // duplicate-pair log: room for every ring slot of 1/8 of the targets, at least 64k rows
        size_t cap = (size_t)n_tiles * tl.n_slots / 8 + 65536;
        WD_TRY(ctx->dup_rows.reserve(cap * 16));
        WD_TRY(ctx->dup_count.reserve(8));
        WD_CUDA(cudaMemsetAsync(ctx->dup_count.p, 0, 8, st));
        ctx->dup_cap = cap;
        a.dup_rows = ctx->dup_rows.as<int32_t>();
        a.dup_count = ctx->dup_count.as<unsigned long long>();
        a.dup_cap = cap;
    } else {
        ctx->dup_cap = 0;
    }
    a.n_head = 0;
    ctx->last_h2d_bytes = 0;
    if (mode == 0 && over_pcie) {
        // Host-mapped tiles: the planes every ring well is read from -- the first positions -- go to HBM
        // by DMA (bandwidth-bound, ~52 GB/s) while the kernel pulls only what the survivors need of the
        // later planes as 32-byte sector reads (request-bound, ~0.3 G requests/s): the two limits of the
        // PCIe path are used side by side, group of tiles after group of tiles.
        int n_head = 2, n_groups = 16;              // profiles/r01_notes.md: sweep on the B200 box
        if (const char *hp = getenv("WELLDUP_HEAD_PLANES")) n_head = atoi(hp);
        if (const char *hg = getenv("WELLDUP_HEAD_GROUPS")) n_groups = atoi(hg);
        n_head = std::max(0, std::min(n_head, std::min(seq_len, 16)));
        for (int k = 0; k < n_tiles; ++k)
            if (ctx->slots[first_slot + k].mapped == nullptr)
                WD_FAIL(WD_E_ARG, "wd_count: host-mapped and staged tile slots cannot share one call");
        if (n_head > 0) {
            n_groups = std::max(1, std::min(n_groups, n_tiles));
            a.n_head = n_head;
            if (!getenv("WELLDUP_STEPS")) {
                a.step0 = n_head;            // first round entirely from HBM
                a.step1 = 1;
            }
            const size_t hstride = ((size_t)ctx->slots[first_slot].n + 255) & ~(size_t)255;
            for (int k = 0; k < n_tiles; ++k) WD_TRY(ctx->slots[first_slot + k].head.reserve(hstride * n_head));
            WD_TRY(upload_descs(ctx, first_slot, n_tiles));       // head pointers may have moved
            while ((int)ctx->copy_events.size() < n_groups + 1) {
                cudaEvent_t ev;
                WD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                ctx->copy_events.push_back(ev);
            }
            // the copies may not overtake earlier work on the compute stream that still reads the head buffers
            WD_CUDA(cudaEventRecord(ctx->copy_events[n_groups], st));
            WD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[n_groups], 0));
            const size_t row = 1 + 2 * (size_t)L;
            for (int g = 0; g < n_groups; ++g) {
                const int t0 = (int)((long long)n_tiles * g / n_groups), t1 = (int)((long long)n_tiles * (g + 1) / n_groups);
                for (int k = t0; k < t1; ++k) {
                    TileSlot &s = ctx->slots[first_slot + k];
                    for (int j = 0; j < n_head; ++j) {
                        const int pl = order[j];
                        const size_t bytes = s.kind[pl] == WD_PLANE_BCL ? (size_t)s.n_block[pl] : ((size_t)s.n_block[pl] + 1) / 2;
                        WD_CUDA(cudaMemcpyAsync(s.head.as<uint8_t>() + (size_t)j * hstride, s.mapped_host + (size_t)pl * s.stride,
                                                bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
                        ctx->last_h2d_bytes += bytes;
                    }
                }
                WD_CUDA(cudaEventRecord(ctx->copy_events[g], ctx->copy_stream));
                WD_CUDA(cudaStreamWaitEvent(st, ctx->copy_events[g], 0));
                CountArgs ag = a;
                ag.descs = a.descs + t0;
                ag.counters = a.counters + (size_t)t0 * width;
                if (a.per_target) ag.per_target = a.per_target + (size_t)t0 * tl.t * row;
                switch (words) {
                    case 1: launch_count_w<1>(ctx, ag, t1 - t0, mode, all_bcl); break;
                    case 2: launch_count_w<2>(ctx, ag, t1 - t0, mode, all_bcl); break;
                    case 4: launch_count_w<4>(ctx, ag, t1 - t0, mode, all_bcl); break;
                    case 8: launch_count_w<8>(ctx, ag, t1 - t0, mode, all_bcl); break;
                    default: launch_count_w<16>(ctx, ag, t1 - t0, mode, all_bcl); break;
                }
            }
        }
    }
    if (a.n_head == 0) {
        switch (words) {
            case 1: launch_count_w<1>(ctx, a, n_tiles, mode, all_bcl); break;
            case 2: launch_count_w<2>(ctx, a, n_tiles, mode, all_bcl); break;
            case 4: launch_count_w<4>(ctx, a, n_tiles, mode, all_bcl); break;
            case 8: launch_count_w<8>(ctx, a, n_tiles, mode, all_bcl); break;
            default: launch_count_w<16>(ctx, a, n_tiles, mode, all_bcl); break;
        }
    }
    WD_CUDA(cudaGetLastError());
    ctx->last_tiles = n_tiles;
    ctx->last_levels = L;
    ctx->last_t = tl.t;
    ctx->last_first_slot = first_slot;
    ctx->last_per_target = want_per_target != 0;
    return WD_OK;
}

int publish_counters(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles, int n_rows_total) {
    if (n_tiles != ctx->last_tiles) WD_FAIL(WD_E_ARG, "wd_publish_counters: last wd_count covered %d tiles, not %d", ctx->last_tiles, n_tiles);
    const int width = 1 + 5 * ctx->last_levels;
    for (int i = 0; i < n_tiles; ++i)
        if (tile_row[i] >= n_rows_total || lane_row[i] >= n_rows_total)
            WD_FAIL(WD_E_ARG, "wd_publish_counters: row index out of range");
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)n_rows_total * width;
    WD_TRY(ctx->publish.reserve(n * 8 + (size_t)n_tiles * 8));
    int32_t *rows = reinterpret_cast<int32_t *>(ctx->publish.as<unsigned long long>() + n);
    WD_CUDA(cudaMemsetAsync(ctx->publish.p, 0, n * 8, st));
    WD_CUDA(cudaMemcpyAsync(rows, tile_row, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(rows + n_tiles, lane_row, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    const int total = n_tiles * width;
    publish_kernel<<<(total + 255) / 256, 256, 0, st>>>(ctx->counters.as<unsigned long long>(), rows, rows + n_tiles,
                                                         n_tiles, width, ctx->publish.as<unsigned long long>());
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    ctx->publish_n = n;
    return WD_OK;
}

// ============================================================================
// Exhaustive mode (BASELINE config 3): every well is a target
// ============================================================================
// No target list exists: a warp takes one centre, walks the three grid rows of
// stage 1 around it (contiguous runs of {x, y, well} records), keeps the wells
// that fall into rings 1..levels under the reference's distance and index
// window rules (prepare_cluster_indexes.py:19,52-67) in a per-warp list in
// shared memory, and then compares the centre's packed words with theirs, 32
// ring wells at a time.  Ring sizes (the LENGTH of the reference) fall out of
// the same walk; a centre with an empty ring is the reference's RuntimeError
// (:70-76) whether or not it passes the filter.
constexpr int EXH_WARPS = 8;
constexpr int EXH_CAP = 256;                       // ring wells kept per centre (a hex lattice has 90)
__constant__ int c_exh_d2[6] = {1, 484, 1764, 3844, 6724, 10404};   // MAX_DISTS^2, as in wd_stage1.cu

struct ExhArgs {
    const int *px, *py;
    const uint32_t *cell_start;
    const int4 *cell_wells;
    const uint8_t *filter;
    const uint64_t *packed;              // [n][W][4]
    unsigned long long *counters;        // [1 + 5 * levels]
    uint32_t *first_empty;               // min over (centre * levels + level) with an empty ring
    uint32_t *overflow;                  // a centre had more than EXH_CAP ring wells
    uint32_t n;
    int levels, len, e, hamming;
    int min_x, min_y, grid_w, grid_h;
    uint32_t wlo, whi;
};

template <int W>
__global__ void __launch_bounds__(EXH_WARPS * 32, W == 1 ? 5 : 1)
exhaustive_count_kernel(ExhArgs a) {
    __shared__ uint32_t s_cnt[1 + 5 * 5];
    __shared__ uint32_t s_list[EXH_WARPS][EXH_CAP];        // well | level << 29
    for (int i = threadIdx.x; i < 1 + 5 * 5; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int L = a.levels;
    const bool ham = a.hamming != 0;
    const bool ham_like = ham || a.e < 2;
    const bool prefilter = a.e >= 0 && a.e < a.len;
    const int d2_max = c_exh_d2[L];
    // consecutive centres to consecutive warps: neighbouring wells share grid rows and packed words in L1/L2
    for (uint32_t c = blockIdx.x * EXH_WARPS + warp; c < a.n; c += gridDim.x * EXH_WARPS) {
        const int cx = __ldg(a.px + c), cy = __ldg(a.py + c);
        // index window [c - wlo, c + whi] (wells are < 2^29, the bounds are clamped to that range)
        const int lo = (int)max((long long)c - (long long)a.wlo, 0ll);
        const int hi = (int)min((long long)c + (long long)a.whi, (long long)0x7fffffff);
        const int x0 = max(cx - RING_RADIUS - a.min_x, 0) >> CELL_SHIFT_X;
        const int x1 = min((cx + RING_RADIUS - a.min_x) >> CELL_SHIFT_X, a.grid_w - 1);
        const int y0 = max(cy - RING_RADIUS - a.min_y, 0) >> CELL_SHIFT_Y;       // two or three grid rows
        const int y1 = min((cy + RING_RADIUS - a.min_y) >> CELL_SHIFT_Y, a.grid_h - 1);
        uint32_t n_ring = 0;
        for (int yy = y0; yy <= y1; ++yy) {
            const uint32_t rs = __ldg(a.cell_start + (uint32_t)yy * a.grid_w + x0);
            const uint32_t re = __ldg(a.cell_start + (uint32_t)yy * a.grid_w + x1 + 1);
            for (uint32_t base = rs; base < re; base += 32) {
                const uint32_t i = base + lane;
                bool in = false;
                uint32_t ent = 0;
                if (i < re) {
                    const int4 w = __ldg(a.cell_wells + i);
                    const int dx = w.x - cx, dy = w.y - cy;         // |dx| < 256 + 32, |dy| < 256: no overflow
                    const int d2 = dx * dx + dy * dy;
                    // MAX[l] < dist <= MAX[l+1]  <=>  MAX[l]^2 < d2 <= MAX[l+1]^2
                    const int lvl = (d2 > 484) + (d2 > 1764) + (d2 > 3844) + (d2 > 6724);
                    in = d2 > 1 && d2 <= d2_max && w.z >= lo && w.z <= hi;
                    ent = (uint32_t)w.z | ((uint32_t)lvl << 29);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const uint32_t pos = n_ring + __popc(m & lt_mask);
                    if (pos < EXH_CAP) s_list[warp][pos] = ent;
                }
                n_ring += __popc(m);
            }
        }
        if (n_ring > EXH_CAP) {
            if (lane == 0) atomicExch(a.overflow, 1u);
            n_ring = EXH_CAP;
        }
        const bool valid = (__ldg(a.filter + c) & 1u) != 0;      // count_well_duplicates.py:236-237
        __syncwarp();
        PSeq<W> cs;
        if (valid) load_packed<W>(a.packed + (size_t)c * (W * PACK_STRIDE), cs);
        uint32_t dups[5] = {0, 0, 0, 0, 0};
        uint32_t lens[5] = {0, 0, 0, 0, 0};
        for (uint32_t base = 0; base < n_ring; base += 32) {
            const uint32_t i = base + lane;
            bool dup = false;
            int lvl = -1;
            if (i < n_ring) {
                const uint32_t ent = s_list[warp][i];
                lvl = (int)(ent >> 29);
                if (valid) {
                    PSeq<W> b;
                    load_packed<W>(a.packed + (size_t)(ent & 0x1fffffffu) * (W * PACK_STRIDE), b);
                    if (!(prefilter && head32_rejects<W>(cs, b, a.len, a.e, ham_like)))
                        dup = is_duplicate<W>(cs, b, a.len, a.e, ham);
                }
            }
#pragma unroll
            for (int l = 0; l < 5; ++l) {
                lens[l] += __popc(__ballot_sync(0xffffffffu, lvl == l));
                dups[l] += __popc(__ballot_sync(0xffffffffu, dup && lvl == l));
            }
        }
        // a ring without wells is the reference's RuntimeError, pass-filter centre or not
#pragma unroll
        for (int l = 0; l < 5; ++l)
            if (l < L && lens[l] == 0 && lane == 0) atomicMin(a.first_empty, c * (uint32_t)L + l);
        if (valid) {
            // the sums of output_writer (count_well_duplicates.py:77-106)
            uint32_t hit_mask = 0;
#pragma unroll
            for (int l = 0; l < 5; ++l)
                if (l < L && dups[l]) hit_mask |= 1u << l;
            if (lane == 0) atomicAdd(&s_cnt[0], 1u);
#pragma unroll
            for (int l = 0; l < 5; ++l) {
                if (l < L && lane == l) {
                    uint32_t *cc = s_cnt + 1 + 5 * l;
                    atomicAdd(cc + 0, lens[l]);
                    if (dups[l]) {
                        atomicAdd(cc + 1, dups[l]);
                        atomicAdd(cc + 2, 1u);
                    }
                    if (hit_mask & ((2u << l) - 1u)) atomicAdd(cc + 3, 1u);
                    if (hit_mask >> l) atomicAdd(cc + 4, 1u);
                }
            }
        }
        __syncwarp();
    }
    flush_counters(s_cnt, a.counters, 1 + 5 * L);
}

int count_exhaustive(wd_ctx *ctx, int slot, const int32_t *order, int seq_len, int levels, uint32_t wlo, uint32_t whi,
                     int e, int hamming, int64_t *tile_counters) {
    TileSlot &s = ctx->slots[slot];
    if (levels < 1 || levels > 5) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: levels must be 1..5 (MAX_DISTS defines 5 rings)");
    if (ctx->n_locs == 0) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: call wd_locs_load first");
    if (ctx->n_locs != s.n)
        WD_FAIL(WD_E_ASSERT, "wd_count_exhaustive: the .locs file holds %u wells, the tile %u", ctx->n_locs, s.n);
    if (s.n >= (1u << 29)) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: at most 2^29 wells per tile");
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, slot, 1, order, seq_len, &all_bcl, &any_excl));
    if (any_excl) WD_TRY(filter_rank(ctx, &slot, 1));
    WD_TRY(upload_descs(ctx, slot, 1));
    cudaStream_t st = ctx->stream;
    const int words = words_for(seq_len);
    const size_t width = 1 + 5 * (size_t)levels;
    WD_TRY(ctx->x_packed.reserve((size_t)s.n * words * PACK_STRIDE * 8));
    WD_TRY(ctx->x_counts.reserve(width * 8 + 8));
    WD_CUDA(cudaMemsetAsync(ctx->x_counts.p, 0, width * 8 + 8, st));
    uint32_t *flags = reinterpret_cast<uint32_t *>(ctx->x_counts.as<unsigned long long>() + width);
    WD_CUDA(cudaMemsetAsync(flags, 0xff, 4, st));
    // dense pass: every well packed once, in index order (coalesced plane reads)
    launch_gather_any(ctx, words, all_bcl, ctx->descs.as<TileDesc>(), nullptr, s.n, 1, seq_len, ctx->x_packed.as<uint64_t>());
    ExhArgs a;
    a.px = ctx->px.as<int>(); a.py = ctx->py.as<int>();
    a.cell_start = ctx->cell_start.as<uint32_t>(); a.cell_wells = ctx->cell_wells.as<int4>();
    a.filter = s.mapped_filter ? s.mapped_filter : s.filter.as<uint8_t>();
    a.packed = ctx->x_packed.as<uint64_t>();
    a.counters = ctx->x_counts.as<unsigned long long>();
    a.first_empty = flags; a.overflow = flags + 1;
    a.n = s.n; a.levels = levels; a.len = seq_len; a.e = e; a.hamming = hamming;
    a.min_x = ctx->min_x; a.min_y = ctx->min_y; a.grid_w = ctx->grid_w; a.grid_h = ctx->grid_h;
    a.wlo = wlo; a.whi = whi;
    const unsigned blocks = (unsigned)std::min<size_t>(((size_t)s.n + EXH_WARPS - 1) / EXH_WARPS, (size_t)ctx->sm_count * 8);
    switch (words) {
        case 1: exhaustive_count_kernel<1><<<blocks, EXH_WARPS * 32, 0, st>>>(a); break;
        case 2: exhaustive_count_kernel<2><<<blocks, EXH_WARPS * 32, 0, st>>>(a); break;
        case 4: exhaustive_count_kernel<4><<<blocks, EXH_WARPS * 32, 0, st>>>(a); break;
        case 8: exhaustive_count_kernel<8><<<blocks, EXH_WARPS * 32, 0, st>>>(a); break;
        default: exhaustive_count_kernel<16><<<blocks, EXH_WARPS * 32, 0, st>>>(a); break;
    }
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h(width + 1);
    WD_CUDA(cudaMemcpyAsync(h.data(), ctx->x_counts.p, (width + 1) * 8, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    uint32_t fl[2];
    memcpy(fl, &h[width], 8);
    if (fl[1]) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: a well has more than %d ring wells; the grid is denser than supported", EXH_CAP);
    if (fl[0] != UINT32_MAX)
        WD_FAIL(WD_E_RUNTIME, "Got no wells for cluster %u level %u", fl[0] / levels, fl[0] % levels);
    for (size_t i = 0; i < width; ++i) tile_counters[i] = (int64_t)h[i];
    return WD_OK;
}

}  // namespace wd
