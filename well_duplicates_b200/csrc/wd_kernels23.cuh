// Stage 2/3 kernel templates (K4/K5 gather-decode, K6 compare + count, the fused
// production kernel) and their launchers.  The templates are instantiated once per
// word count W in wd_inst23_w*.cu, so that the five flavours compile in parallel;
// wd_stage23.cu holds the host side and only sees `extern template` declarations.
#pragma once
#include "wd_common.cuh"
#include "wd_pack.cuh"
#include "wd_seq.cuh"

namespace wd {

constexpr int MAX_ORDER = WD_MAX_SEQ_LEN;

__device__ __forceinline__ int pf_rank(const TileDesc &d, uint32_t well) {
    const uint64_t m = __ldg(d.pfmask + (well >> 6));
    const int b = well & 63;
    if (!((m >> b) & 1ull)) return -1;
    return (int)(__ldg(d.pfrank + (well >> 6)) + (uint32_t)__popcll(m & ((1ull << b) - 1ull)));
}

// ============================================================================
// K4 / K5: gather-decode one well into packed words
// ============================================================================
// s_off[p]  = byte offset of the plane that supplies sequence position p
// s_kind[p] = WD_PLANE_* of that plane (same for every tile of a launch)
// one base call -> raw code: 0 = no-call, otherwise base = code & 3
// A scattered 1-byte read of a plane.  A plain load makes L2 fill the whole 128-byte line from DRAM (4 sectors
// for 1 useful byte); ld.global.nc.L2::64B is the smallest fill this part offers (profiles/
// r02_fetch_granularity_micro.txt) and halves the DRAM bytes of the gather -- but the gather is bound by the number
// of requests to DRAM, not by their size, and whole lines serve neighbouring targets: 64-byte fills were measured
// 4 % SLOWER (0.398 vs 0.381 ms per 96-tile launch, profiles/r02_notes.md).  Kept as a build option.  (For planes in
// host memory the prefetch size changes nothing: sector pulls across PCIe stay 32-byte requests, 0.35 G/s.)
#ifndef WD_PLANE_LD_L2_64B
#define WD_PLANE_LD_L2_64B 0
#endif
__device__ __forceinline__ uint32_t ld_plane_u8(const uint8_t *p) {
#if WD_PLANE_LD_L2_64B
    uint32_t v;
    asm("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

template <bool ALL_BCL>
__device__ __forceinline__ uint32_t load_call(const TileDesc &d, uint32_t well, int rank, unsigned long long off,
                                              int kind) {
    if (ALL_BCL || kind == WD_PLANE_BCL) return ld_plane_u8(d.planes + off + well);
    const int wi = kind == WD_PLANE_CBCL_EXCL ? rank : (int)well;
    if (wi < 0) return 0u;                        // not PF: the block has no entry for it -> N
    const uint32_t byte = ld_plane_u8(d.planes + off + ((uint32_t)wi >> 1));
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

// a 32-bit group starting at any position (may straddle a 64-bit word)
template <int W>
__device__ __forceinline__ void pseq_or_bits(PSeq<W> &q, int p, uint32_t glo, uint32_t ghi, uint32_t gnn) {
    const int w = p >> 6, sh = p & 63;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == w) {
            q.lo[i] |= (uint64_t)glo << sh;
            q.hi[i] |= (uint64_t)ghi << sh;
            q.nn[i] |= (uint64_t)gnn << sh;
        }
        if (i > 0 && i == w + 1 && sh > 32) {
            q.lo[i] |= (uint64_t)glo >> (64 - sh);
            q.hi[i] |= (uint64_t)ghi >> (64 - sh);
            q.nn[i] |= (uint64_t)gnn >> (64 - sh);
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

// dense flavour (exhaustive mode): every well of the tile in index order, BCL
// planes only.  One thread packs four consecutive wells from 32-bit plane loads
// (a warp reads 128 contiguous bytes of a plane per instruction), 16 loads in
// flight.  The four calls of a word are decoded together, byte-sliced:
//   bit0 of the bases   v & 0x01010101
//   bit1                (v >> 1) & 0x01010101
//   no-call (byte == 0) ~(((v & 0x7f7f7f7f) + 0x7f7f7f7f) | v) >> 7 & 0x01010101   (exact, no cross-byte carry)
// and cycle j of a group of eight lands at bit j of each byte by one multiply-add
// (the integer FMA pipe: the logic pipe is the busy one); a 4 x 4 byte transpose
// (PRMT) then turns four groups into one 32-bit run of symbols per well.
__device__ __forceinline__ void transpose4x4_bytes(const uint32_t (&a)[4], uint32_t (&out)[4]) {
    const uint32_t t0 = __byte_perm(a[0], a[1], 0x5140), t1 = __byte_perm(a[0], a[1], 0x7362);
    const uint32_t u0 = __byte_perm(a[2], a[3], 0x5140), u1 = __byte_perm(a[2], a[3], 0x7362);
    out[0] = __byte_perm(t0, u0, 0x5410);
    out[1] = __byte_perm(t0, u0, 0x7632);
    out[2] = __byte_perm(t1, u1, 0x5410);
    out[3] = __byte_perm(t1, u1, 0x7632);
}

template <int W>
__global__ void __launch_bounds__(256)
dense_pack_bcl_kernel(const TileDesc *__restrict__ descs, uint32_t n, const unsigned long long *__restrict__ g_off,
                      const uint8_t *__restrict__ g_kind, int len, uint64_t *__restrict__ packed) {
    __shared__ unsigned long long s_off[MAX_ORDER];
    __shared__ uint8_t s_kind[MAX_ORDER];
    load_order(g_off, g_kind, len, s_off, s_kind);
    const uint32_t w0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (w0 >= n) return;
    const TileDesc d = descs[blockIdx.y];
    const uint8_t *src = d.planes + w0;              // planes are padded to a multiple of 256 bytes: 4-byte loads stay inside
    uint64_t *dst = packed + ((size_t)blockIdx.y * n + w0) * (size_t)(W * PACK_STRIDE);
    uint32_t pf = 0;                                 // bit 8 i = PF of well i
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (w0 + i < n) pf |= (uint32_t)(__ldg(d.filter + w0 + i) & 1u) << (8 * i);
#pragma unroll 1
    for (int w = 0; w < W; ++w) {
        uint32_t lo[2][4], hi[2][4], nn[2][4];       // [half of the 64-symbol word][well]
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t gl[4], gh[4], gn[4];            // one group of eight cycles each, byte i = well i
#pragma unroll
            for (int g2 = 0; g2 < 4; g2 += 2) {
                const int p0 = 64 * w + 32 * half + 8 * g2;
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = 0x04040404u;              // beyond the sequence: no bit in any plane
                    if (p0 + j < len) v[j] = __ldg(reinterpret_cast<const uint32_t *>(src + s_off[p0 + j]));
                }
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    uint32_t al = 0, ah = 0, an = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t x = v[8 * gg + j];
                        const uint32_t z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x);
                        al += (x & 0x01010101u) * (1u << j);
                        ah += ((x >> 1) & 0x01010101u) * (1u << j);
                        an += ((z >> 7) & 0x01010101u) * (1u << j);
                    }
                    gl[g2 + gg] = al; gh[g2 + gg] = ah; gn[g2 + gg] = an;
                }
            }
            transpose4x4_bytes(gl, lo[half]);
            transpose4x4_bytes(gh, hi[half]);
            transpose4x4_bytes(gn, nn[half]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (w0 + i < n) {
                ulonglong2 *q = reinterpret_cast<ulonglong2 *>(dst + ((size_t)i * W + w) * PACK_STRIDE);
                q[0] = make_ulonglong2((uint64_t)lo[0][i] | ((uint64_t)lo[1][i] << 32), (uint64_t)hi[0][i] | ((uint64_t)hi[1][i] << 32));
                q[1] = make_ulonglong2((uint64_t)nn[0][i] | ((uint64_t)nn[1][i] << 32), w == 0 ? (uint64_t)((pf >> (8 * i)) & 1u) : 0ull);
            }
        }
    }
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
    int32_t *dup_rows;               // may be null: (tile, target, slot, distance)
    unsigned long long *dup_count;
    unsigned long long dup_cap;
    unsigned long long *status;      // [0]: ~(tile << 32 | target) of the first valid target with an empty ring (0: none)
    uint32_t *trace;                 // measurement build of the fused kernel only: one bit per 32-byte sector read,
    uint32_t trace_words;            //   [tile][position][trace_words]
    uint32_t tile_base;              // index, in the caller's batch, of the launch's first tile
    uint32_t t, n_slots;
    int levels, len, e, hamming;
    int step0, step1;                // fused kernel: cycles read per round (first, later), 1..8
    int cchunk;                      // fused kernel: 0 = the centre is read exactly as far as a round needs it; 8, 16, 32:
                                     //   rounded up to a multiple of that (measurement sweeps)
    int n_head;                      // fused kernel: positions 0..n_head-1 are read from the tile's head planes in HBM
    int tpb;                         // fused kernel: targets per CTA (grid.x = ceil(t / tpb))
    int pad_smem;                    // fused kernel (host side only): bytes of dynamic shared memory the launch reserves to
                                     //   cap the CTAs resident per SM
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
            const uint32_t ring = __ldg(a.level_len + (size_t)t * L + l);
            // count_well_duplicates.py:249 asserts that a ring holds wells -- for targets that get this far
            if (ring == 0) atomicMax(a.status, ~(((unsigned long long)(a.tile_base + tile) << 32) | t));
            atomicAdd(c + 0, ring);
            if (dups[l]) {
                atomicAdd(c + 1, dups[l]);
                atomicAdd(c + 2, 1u);
            }
            if (hit_mask & ((2u << l) - 1u)) atomicAdd(c + 3, 1u);
            if (hit_mask >> l) atomicAdd(c + 4, 1u);
        }
    }
}

// one row of the duplicate-pair log (count_well_duplicates.py:258-262); rows beyond the buffer are counted only
__device__ __forceinline__ void log_dup_row(const CountArgs &a, uint32_t tile, uint32_t t, uint32_t slot, int dist) {
    const unsigned long long pos = atomicAdd(a.dup_count, 1ull);
    if (pos < a.dup_cap)
        *reinterpret_cast<int4 *>(a.dup_rows + pos * 4) = make_int4((int)(a.tile_base + tile), (int)t, (int)slot, dist);
}

// The fused kernel's version.  Wells and Targets are the same sums for every target without a duplicate (98 % of
// them), so they are kept in ONE register per lane across the CTA's targets -- lane l < L adds the size of ring l,
// lane 31 counts the targets -- and reach shared memory once per warp (flush_warp_tallies); only a target with a hit
// goes through the shared-memory counters for Dups / Hit / AccO / AccI.
template <int LMAX>
__device__ __forceinline__ void finish_target_fused(const CountArgs &a, uint32_t tile, uint32_t t, int lane, bool valid,
                                                    const uint32_t *dups, uint32_t *s_cnt, uint32_t &tally) {
    const int L = a.levels;
    if (a.per_target != nullptr) {
        const int row = 1 + 2 * L;
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
    if (lane < L) {
        const uint32_t ring = __ldg(a.level_len + (size_t)t * L + lane);
        // count_well_duplicates.py:249 asserts that a ring holds wells -- for targets that get this far
        if (ring == 0) atomicMax(a.status, ~(((unsigned long long)(a.tile_base + tile) << 32) | t));
        tally += ring;
    } else if (lane == 31) {
        tally += 1u;
    }
    uint32_t hit_mask = 0;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
        if (l < L && dups[l]) hit_mask |= 1u << l;
    if (hit_mask == 0) return;
    // AccO: a hit at this level or further in; AccI: at this level or further out (count_well_duplicates.py:77-89)
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
        if (l < L && lane == l) {
            uint32_t *c = s_cnt + 1 + 5 * l;
            if (dups[l]) {
                atomicAdd(c + 1, dups[l]);
                atomicAdd(c + 2, 1u);
            }
            if (hit_mask & ((2u << l) - 1u)) atomicAdd(c + 3, 1u);
            if (hit_mask >> l) atomicAdd(c + 4, 1u);
        }
    }
}

__device__ __forceinline__ void flush_warp_tallies(int levels, int lane, uint32_t tally, uint32_t *s_cnt) {
    if (lane < levels) atomicAdd(s_cnt + 1 + 5 * lane, tally);
    else if (lane == 31) atomicAdd(s_cnt, tally);
}

template <int W>
__device__ __forceinline__ void log_dup(const CountArgs &a, uint32_t tile, uint32_t t, uint32_t slot,
                                        const PSeq<W> &c, const PSeq<W> &b) {
    log_dup_row(a, tile, t, slot, exact_distance<W>(c, b, a.len, a.hamming != 0));
}

__device__ __forceinline__ void flush_counters(uint32_t *s_cnt, unsigned long long *dst, int n) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(dst + i, (unsigned long long)s_cnt[i]);
}

constexpr int CNT_WARPS = 8;

// once per tile: an excluded CBCL block holds exactly the wells that pass the filter (cbcl_read.py:130-131)
__device__ __forceinline__ void check_excluded_total(const CountArgs &a, const TileDesc &d, uint32_t tile) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && (d.flags & 2u) && __ldg(d.pfrank + (d.n + 63) / 64) != d.excl_expect)
        atomicMax(a.status + 1, ~(unsigned long long)(a.tile_base + tile));
}

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
    check_excluded_total(a, a.descs[tile], tile);
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
                // most passes end without a duplicate: one vote decides whether the per-level tallies are needed
                if (__any_sync(0xffffffffu, dup)) {
#pragma unroll
                    for (int l = 0; l < LMAX; ++l)
                        if (l < a.levels) dups[l] += __popc(__ballot_sync(0xffffffffu, dup && lvl == l + 1));
                }
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
// Targets per CTA (CountArgs::tpb): its 8 warps pull them from a shared counter.  count_run chooses the number per
// launch (wd_stage23.cu, measurements in profiles/r02_notes.md); wd_set_tuning overrides it.

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

// Shared-memory reads through 32-bit shared-window addresses.  Indexing a __shared__ array through a generic
// pointer makes the compiler rebuild the window base (S2R SR_CgaCtaId + LEA) next to every load once registers are
// tight -- five extra instructions per plane offset in the inner loop of the fused kernel.
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
    unsigned long long v;
    asm("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");      // the Eq table is rewritten per target
    return v;
}
// a value the compiler may not recompute: it has to stay in a register
__device__ __forceinline__ uint32_t keep_in_register(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// measurement build: mark the 32-byte sector of the plane of position `pos` that load_call reads for this well
template <bool ALL_BCL>
__device__ __forceinline__ void trace_call(const CountArgs &a, uint32_t tile, int pos, uint32_t well, int rank, int kind) {
    uint32_t byte = well;
    if (!(ALL_BCL || kind == WD_PLANE_BCL)) {
        const int wi = kind == WD_PLANE_CBCL_EXCL ? rank : (int)well;
        if (wi < 0) return;
        byte = (uint32_t)wi >> 1;
    }
    const uint32_t sec = byte >> 5;              // planes start on 256-byte boundaries
    atomicOr(a.trace + ((size_t)tile * a.len + pos) * a.trace_words + (sec >> 5), 1u << (sec & 31));
}

// One round of a ring well: its n (<= NMAX) calls raw[0..n) at positions p .. p+n-1 are fed to the distance test;
// false = the prefix proves dist > e.  eqtab: this warp's Eq masks of word 0 (one per symbol), for the rounds
// inside the first 32 rows.
template <int W, int NMAX>
__device__ __forceinline__ bool ring_compute(const uint32_t (&raw)[8], uint32_t eqtab, const PSeq<W> &c, int known_c,
                                             int len, int p, int n, int k, int e, bool ham_like, PrefixDP<W> &dp, int &mism) {
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
    if (pdp_round_in_word0(p, n, k)) {
        // the first round of every well: only word 0 of the programme is active (wd_seq.cuh), and the Eq mask of
        // a symbol is one shared-memory read instead of seven logic operations
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n) pdp_step_word0_eq<W>(dp, lds_u32(eqtab + 4u * call_symbol(raw[j])));
        return pdp_band_min_word0<W>(dp, len, p + n, k) <= e;
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
        if (j < n) pdp_step<W>(dp, c, known_c, len, p + j, k, call_symbol(raw[j]));
    return pdp_band_min<W>(dp, len, p + n, k) <= e;
}

// resident CTAs per SM the register allocation aims at for the common flavour (W = 1): 8 x 8 warps = every warp
// slot of the SM at 32 registers, no spills.  The gather is bound by requests in flight: 0.368 ms per 96-tile
// launch against 0.385 at 6 CTAs (38 registers) and 0.415 at 5 (profiles/r02_notes.md).
#ifndef WD_FUSED_MIN_CTAS
#define WD_FUSED_MIN_CTAS 8
#endif

template <int W, int LMAX, bool ALL_BCL, bool TRACE>
__global__ void __launch_bounds__(CNT_WARPS * 32, W == 1 ? WD_FUSED_MIN_CTAS : 1)
fused_count_kernel(CountArgs a) {
    __shared__ unsigned long long s_off[MAX_ORDER];
    __shared__ uint8_t s_kind[MAX_ORDER];
    __shared__ uint32_t s_cnt[1 + 5 * LMAX];
    __shared__ uint32_t s_next;
    __shared__ uint32_t s_eq[CNT_WARPS][8];         // per warp: Eq mask of word 0 for symbols A, C, G, T, N
    for (int i = threadIdx.x; i < 1 + 5 * LMAX; i += blockDim.x) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    load_order(a.g_off, a.g_kind, a.len, s_off, s_kind);
    const int lane = threadIdx.x & 31;
    uint32_t *eqtab_w = s_eq[threadIdx.x >> 5];
    const uint32_t eqtab = keep_in_register((uint32_t)__cvta_generic_to_shared(eqtab_w));
    const uint32_t off_sh = keep_in_register((uint32_t)__cvta_generic_to_shared(s_off));
    const uint32_t kind_sh = ALL_BCL ? 0u : keep_in_register((uint32_t)__cvta_generic_to_shared(s_kind));
    const uint32_t tile = blockIdx.y;
    const TileDesc d = a.descs[tile];
    const int len = a.len, e = a.e;
    if (!ALL_BCL) check_excluded_total(a, d, tile);
    if (a.n_head) {
        // this CTA serves one tile: point its first positions at the tile's head planes
        for (int i = threadIdx.x; i < a.n_head; i += blockDim.x) s_off[i] = d.head_delta + (unsigned long long)i * d.head_stride;
        __syncthreads();
    }
    // Levenshtein <= 1 <=> Hamming <= 1 on equal lengths (an indel pair costs 2)
    const bool ham_like = a.hamming != 0 || e < 2;
    const int k = ham_like ? 0 : (e >> 1);
    // no pair / every pair is a duplicate -- but the log wants the distance of every duplicate
    const bool read_nothing = e < 0 || (e >= len && a.dup_rows == nullptr);
    const uint32_t t_begin = blockIdx.x * (uint32_t)a.tpb;
    const uint32_t t_end = min(t_begin + (uint32_t)a.tpb, a.t);
    uint32_t tally = 0;                                // Wells per level (lanes 0..L-1) and Targets (lane 31) of this warp
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
            int known_c = 0, eq_known = -1;
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
                    // ---- ring wells, lane = well: the round's calls, all loads in flight together ----------
                    uint32_t raw[8];
                    if (n == 8) {
                        // the first round of the resident schedule: one predicate for the eight loads
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            raw[j] = 0u;
                            if (alive) {
                                const int kind = ALL_BCL ? 0 : (int)lds_u8(kind_sh + p + j);
                                raw[j] = load_call<ALL_BCL>(d, well, rank, lds_u64(off_sh + 8u * (p + j)), kind);
                                if (TRACE) trace_call<ALL_BCL>(a, tile, p + j, well, rank, kind);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            raw[j] = 0u;
                            if (alive && j < n) {
                                const int kind = ALL_BCL ? 0 : (int)lds_u8(kind_sh + p + j);
                                raw[j] = load_call<ALL_BCL>(d, well, rank, lds_u64(off_sh + 8u * (p + j)), kind);
                                if (TRACE) trace_call<ALL_BCL>(a, tile, p + j, well, rank, kind);
                            }
                        }
                    }
                    // ---- centre, lane = cycle: exactly as far as this round looks ahead (k symbols past the ring
                    // wells; a.cchunk > 0 rounds that up), in flight together with the ring loads ------------------
                    int need = min(len, p + n + k);
                    if (a.cchunk > 0) need = min(len, (need + a.cchunk - 1) & ~(a.cchunk - 1));      // 8, 16 or 32
                    while (known_c < need) {
                        const int upto = min(need, known_c + 32);
                        const int q = known_c + lane;
                        uint32_t sym = 0u;
                        if (q < upto) {
                            const int kind = ALL_BCL ? 0 : (int)lds_u8(kind_sh + q);
                            sym = call_symbol(load_call<ALL_BCL>(d, centre, crank, lds_u64(off_sh + 8u * q), kind));
                            if (TRACE) trace_call<ALL_BCL>(a, tile, q, centre, crank, kind);
                        }
                        const uint32_t glo = __ballot_sync(0xffffffffu, sym & 1u);
                        const uint32_t ghi = __ballot_sync(0xffffffffu, sym & 2u);
                        const uint32_t gnn = __ballot_sync(0xffffffffu, sym & 4u);
                        pseq_or_bits<W>(c, known_c, glo, ghi, gnn);
                        known_c = upto;
                    }
                    if (!ham_like && eq_known != known_c && pdp_round_in_word0(p, n, k)) {
                        // Eq masks of word 0 against what is known of the centre, one per symbol
                        __syncwarp();
                        if (lane < 5) {
                            const uint32_t tlo = (lane & 1) ? ~0u : 0u, thi = (lane & 2) ? ~0u : 0u, tn = (lane & 4) ? ~0u : 0u;
                            eqtab_w[lane] = ~(((uint32_t)c.lo[0] ^ tlo) | ((uint32_t)c.hi[0] ^ thi) | ((uint32_t)c.nn[0] ^ tn)) &
                                          len_mask32(known_c, 0);
                        }
                        __syncwarp();
                        eq_known = known_c;
                    }
                    if (alive) {
                        // three sizes of round are compiled (the schedules that win use 8 + 2 or 8 + 4): a
                        // smaller kernel, fewer instruction-cache misses
                        if (n > 4) alive = ring_compute<W, 8>(raw, eqtab, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                        else if (n > 2) alive = ring_compute<W, 4>(raw, eqtab, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                        else alive = ring_compute<W, 2>(raw, eqtab, c, known_c, len, p, n, k, e, ham_like, dp, mism);
                    }
                    p += n;
                }
                const bool dup = alive;            // survived to p == len: dist <= e
                // the log: a survivor has been fed every symbol, so the programme holds the exact distance
                // (at p == len the band minimum is D[len][len]; Lev == Ham where Ham <= 1)
                if (dup && a.dup_rows != nullptr)
                    log_dup_row(a, tile, t, s, ham_like ? mism : pdp_band_min<W>(dp, len, len, k));
                // most passes end without a duplicate: one vote decides whether the per-level tallies are needed
                if (__any_sync(0xffffffffu, dup)) {
#pragma unroll
                    for (int l = 0; l < LMAX; ++l)
                        if (l < a.levels) dups[l] += __popc(__ballot_sync(0xffffffffu, dup && lvl == l + 1));
                }
            }
        }
        finish_target_fused<LMAX>(a, tile, t, lane, valid, dups, s_cnt, tally);
    }
    flush_warp_tallies(a.levels, lane, tally, s_cnt);
    flush_counters(s_cnt, a.counters + (size_t)tile * (1 + 5 * a.levels), 1 + 5 * a.levels);
}


// ============================================================================
// launchers (one explicit instantiation per W)
// ============================================================================
template <int W, bool ALL_BCL>
void launch_gather(wd_ctx *ctx, const TileDesc *descs, const uint32_t *slot_well, uint32_t n_slots, int n_tiles,
                          int len, uint64_t *packed) {
    const unsigned long long *g_off = ctx->order_dev.as<unsigned long long>();
    const uint8_t *g_kind = ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8;
    if (ALL_BCL && slot_well == nullptr) {
        dim3 grid((n_slots + 1023) / 1024, n_tiles);
        dense_pack_bcl_kernel<W><<<grid, 256, 0, ctx->stream>>>(descs, n_slots, g_off, g_kind, len, packed);
    } else {
        dim3 grid((n_slots + 255) / 256, n_tiles);
        gather_pack_kernel<W, ALL_BCL><<<grid, 256, 0, ctx->stream>>>(descs, slot_well, n_slots, g_off, g_kind, len, packed);
    }
    ctx->launches++;
}

template <int W>
void launch_gather_w(wd_ctx *ctx, bool all_bcl, const TileDesc *descs, const uint32_t *slot_well,
                            uint32_t n_slots, int n_tiles, int len, uint64_t *packed) {
    if (all_bcl) launch_gather<W, true>(ctx, descs, slot_well, n_slots, n_tiles, len, packed);
    else launch_gather<W, false>(ctx, descs, slot_well, n_slots, n_tiles, len, packed);
}

template <int W>
void launch_unpack_w(wd_ctx *ctx, const uint64_t *packed, uint32_t n_idx, int len, uint8_t *codes, uint8_t *pf) {
    unpack_codes_kernel<W><<<(n_idx + 255) / 256, 256, 0, ctx->stream>>>(packed, n_idx, len, codes, pf);
    ctx->launches++;
}

// a launch may reserve most of the SM's shared memory (CountArgs::pad_smem, a measurement knob: asked per launch)
template <class K>
void allow_dynamic_smem(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

template <int W, int LMAX>
void launch_count(wd_ctx *ctx, const CountArgs &a, int n_tiles, int mode, bool all_bcl) {
    dim3 grid((a.t + CNT_WARPS - 1) / CNT_WARPS, n_tiles);
    dim3 fgrid((a.t + a.tpb - 1) / a.tpb, n_tiles);
    const size_t pad = (size_t)std::max(0, a.pad_smem);
    if (mode == 1) {
        compare_count_kernel<W, LMAX><<<grid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
    } else if (a.trace != nullptr) {
        // measurement build (wd_count_trace_sectors): one flavour only, the caller has checked that it applies
        if constexpr (W == 1 && LMAX == 5) {
            if (all_bcl) fused_count_kernel<1, 5, true, true><<<fgrid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
            else fused_count_kernel<1, 5, false, true><<<fgrid, CNT_WARPS * 32, 0, ctx->stream>>>(a);
        }
    } else if (all_bcl) {
        if (pad > 48 * 1024) allow_dynamic_smem(fused_count_kernel<W, LMAX, true, false>);
        fused_count_kernel<W, LMAX, true, false><<<fgrid, CNT_WARPS * 32, pad, ctx->stream>>>(a);
    } else {
        if (pad > 48 * 1024) allow_dynamic_smem(fused_count_kernel<W, LMAX, false, false>);
        fused_count_kernel<W, LMAX, false, false><<<fgrid, CNT_WARPS * 32, pad, ctx->stream>>>(a);
    }
    ctx->launches++;
}

template <int W>
void launch_count_w(wd_ctx *ctx, const CountArgs &a, int n_tiles, int mode, bool all_bcl) {
    if (a.levels <= 5) launch_count<W, 5>(ctx, a, n_tiles, mode, all_bcl);
    else launch_count<W, WD_MAX_LEVELS>(ctx, a, n_tiles, mode, all_bcl);
}

#define WD_FOR_EACH_W(X) X(1) X(2) X(4) X(8) X(16)
#define WD_DECLARE_W(W)                                                                                          \
    extern template void launch_gather_w<W>(wd_ctx *, bool, const TileDesc *, const uint32_t *, uint32_t, int, int, \
                                            uint64_t *);                                                         \
    extern template void launch_unpack_w<W>(wd_ctx *, const uint64_t *, uint32_t, int, uint8_t *, uint8_t *);     \
    extern template void launch_count_w<W>(wd_ctx *, const CountArgs &, int, int, bool);
#define WD_INSTANTIATE_W(W)                                                                                      \
    template void launch_gather_w<W>(wd_ctx *, bool, const TileDesc *, const uint32_t *, uint32_t, int, int,        \
                                     uint64_t *);                                                                \
    template void launch_unpack_w<W>(wd_ctx *, const uint64_t *, uint32_t, int, uint8_t *, uint8_t *);            \
    template void launch_count_w<W>(wd_ctx *, const CountArgs &, int, int, bool);

}  // namespace wd
