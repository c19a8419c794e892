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
#include "wd_kernels23.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace wd {

WD_FOR_EACH_W(WD_DECLARE_W)

// ============================================================================
// K3: filter bytes -> PF bit mask + block ranks
// ============================================================================
// A batch of tiles in two launches (round 1: four launches and a memset per tile).  The filter bytes are read from
// HBM; those of a host-mapped tile are brought there by DMA first (51 GB/s) -- reading them in place across PCIe
// from the mask kernel was measured slower (331 vs 301 ms per 704-tile CBCL lane, profiles/r02_notes.md).
struct RankJob {
    const uint8_t *filt;     // n filter bytes, 16-byte aligned (device view)
    uint64_t *mask;          // [nb]   PF bit per well
    uint32_t *rank;          // [nb+1] PF wells before each block of 64; rank[nb] = the tile's PF total
    uint32_t n, nb;
};

// one thread per 64 wells of tile blockIdx.y: mask word + its population count (kept in rank[b + 1] for the scan)
__global__ void __launch_bounds__(256)
filter_mask_kernel(const RankJob *__restrict__ jobs) {
    const RankJob j = jobs[blockIdx.y];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= j.nb) return;
    uint64_t m = 0;
    if ((size_t)b * 64 + 64 <= j.n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(j.filt + (size_t)b * 64);
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
    } else {
        for (uint32_t i = b * 64; i < j.n; ++i) m |= (uint64_t)(__ldg(j.filt + i) & 1u) << (i & 63);     // the tile's last block
    }
    j.mask[b] = m;
    j.rank[b + 1] = (uint32_t)__popcll(m);
}

// one CTA per tile: rank[b] = sum of the counts in front of block b (exclusive scan in place, rank[0] = 0)
__global__ void __launch_bounds__(1024)
filter_rank_kernel(const RankJob *__restrict__ jobs) {
    const RankJob j = jobs[blockIdx.x];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) { s_carry = 0; j.rank[0] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < j.nb; base += 1024) {
        const uint32_t b = base + threadIdx.x;
        const uint32_t v = b < j.nb ? j.rank[b + 1] : 0u;
        uint32_t x = v;                                     // inclusive scan of the chunk
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t before = s_carry + (warp ? s_warp[warp - 1] : 0u);
        if (b < j.nb) j.rank[b + 1] = before + x;           // inclusive at b = exclusive at b + 1
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + x;
        __syncthreads();
    }
}

// the reference's filter_offsets list (parity hook)
__global__ void __launch_bounds__(256)
filter_expand_kernel(TileDesc d, int32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.n) out[i] = pf_rank(d, i);
}

int filter_rank(wd_ctx *ctx, const int *slot_ids, int n) {
    cudaStream_t st = ctx->stream;
    std::vector<RankJob> jobs;
    uint32_t max_nb = 0;
    for (int k = 0; k < n; ++k) {
        TileSlot &s = ctx->slots[slot_ids[k]];
        if (s.rank_valid) continue;
        if (!s.filter_set) WD_FAIL(WD_E_ARG, "tile slot %d has no filter loaded", slot_ids[k]);
        const uint32_t nb = (s.n + 63) / 64;
        if ((size_t)nb * 8 > s.pfmask.cap || ((size_t)nb + 1) * 4 > s.pfrank.cap) WD_CUDA(cudaStreamSynchronize(st));
        WD_TRY(s.pfmask.reserve((size_t)nb * 8));
        WD_TRY(s.pfrank.reserve(((size_t)nb + 1) * 4));
        if (s.mapped_filter) WD_CUDA(cudaMemcpyAsync(s.filter.p, s.mapped_filter_host, s.n, cudaMemcpyHostToDevice, st));
        RankJob j;
        j.filt = s.filter.as<uint8_t>();
        j.mask = s.pfmask.as<uint64_t>();
        j.rank = s.pfrank.as<uint32_t>();
        j.n = s.n;
        j.nb = nb;
        if (((uintptr_t)j.filt & 15u) != 0) WD_FAIL(WD_E_ARG, "tile slot %d: the filter bytes must be 16-byte aligned", slot_ids[k]);
        jobs.push_back(j);
        max_nb = std::max(max_nb, nb);
        s.rank_valid = true;
    }
    if (jobs.empty()) return WD_OK;
    if (jobs.size() * sizeof(RankJob) > ctx->rank_jobs.cap) WD_CUDA(cudaStreamSynchronize(st));
    WD_TRY(ctx->rank_jobs.reserve(jobs.size() * sizeof(RankJob)));
    WD_CUDA(cudaMemcpyAsync(ctx->rank_jobs.p, jobs.data(), jobs.size() * sizeof(RankJob), cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaStreamSynchronize(st));              // `jobs` is pageable and local
    const RankJob *d_jobs = ctx->rank_jobs.as<RankJob>();
    filter_mask_kernel<<<dim3((max_nb + 255) / 256, (unsigned)jobs.size()), 256, 0, st>>>(d_jobs);
    filter_rank_kernel<<<(unsigned)jobs.size(), 1024, 0, st>>>(d_jobs);
    ctx->launches += 2;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

static TileDesc make_desc(const TileSlot &s) {
    TileDesc d;
    memset(&d, 0, sizeof(d));                     // descriptors are compared bytewise (upload_descs)
    d.planes = s.mapped ? s.mapped : s.planes.as<uint8_t>();
    d.stride = s.stride;
    d.filter = s.mapped_filter ? s.mapped_filter : s.filter.as<uint8_t>();
    d.pfmask = s.pfmask.as<uint64_t>();
    d.pfrank = s.pfrank.as<uint32_t>();
    d.kind = s.kind_dev.as<uint8_t>();
    d.n = s.n;
    d.flags = s.rank_valid ? 1u : 0u;
    if (s.has_excl && s.rank_valid) {
        for (int p = 0; p < s.n_planes; ++p)
            if (s.kind[p] == WD_PLANE_CBCL_EXCL) {
                d.excl_expect = s.n_block[p];
                d.flags |= 2u;
                break;
            }
    }
    d.head_stride = s.head_stride;
    d.head_delta = (unsigned long long)(uintptr_t)s.head_ptr - (unsigned long long)(uintptr_t)d.planes;
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
// K7: place tile counters into the all-reduce buffer
// ============================================================================
__global__ void publish_kernel(const unsigned long long *__restrict__ counters, const int32_t *__restrict__ tile_row,
                               const int32_t *__restrict__ lane_row, int n_tiles, int width,
                               unsigned long long *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * width) return;
    const int tile = i / width, k = i % width;
    const unsigned long long v = counters[i];
    if (tile_row[tile] >= 0) out[(size_t)tile_row[tile] * width + k] = v;      // a row belongs to one tile of one rank
    if (lane_row[tile] >= 0 && v) atomicAdd(out + (size_t)lane_row[tile] * width + k, v);
}

// ============================================================================
// duplicate-pair log and the sector trace: small kernels behind wd_dup_pairs_seqs / wd_count_trace_sectors
// ============================================================================
// the two sequences of every logged pair, one byte per symbol (0..3 ACGT, 4 N): codes[row][centre, well][len].
// One thread per symbol: with host-mapped planes every byte is a read across PCIe, and a thread that walked a whole
// sequence would pay those round trips one after the other (2.4 ms for the 2128 pairs of the benchmark lane).
template <bool ALL_BCL>
__global__ void __launch_bounds__(128)
dup_seq_kernel(const TileDesc *__restrict__ descs, const int32_t *__restrict__ rows, unsigned long long n_rows,
               const uint32_t *__restrict__ slot_well, const uint32_t *__restrict__ tgt_off,
               const unsigned long long *__restrict__ g_off, const uint8_t *__restrict__ g_kind, int len,
               uint8_t *__restrict__ codes) {
    const unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (j >= 2 * n_rows * (unsigned long long)len) return;
    const unsigned long long i = j / (unsigned)len;           // (row, centre or well)
    const int p = (int)(j % (unsigned)len);
    const int32_t *r = rows + (i >> 1) * 4;
    const TileDesc d = descs[r[0]];
    const uint32_t well = __ldg(slot_well + ((i & 1) ? (uint32_t)r[2] : __ldg(tgt_off + r[1])));
    int rank = 0;
    if (!ALL_BCL && (d.flags & 1u)) rank = pf_rank(d, well);
    codes[j] = (uint8_t)call_symbol(load_call<ALL_BCL>(d, well, rank, __ldg(g_off + p), ALL_BCL ? 0 : (int)__ldg(g_kind + p)));
}

// sector bitmaps of the measurement build -> per (tile, position): distinct 32-byte sectors, distinct 128-byte lines
__global__ void __launch_bounds__(256)
trace_reduce_kernel(const uint32_t *__restrict__ trace, uint32_t words, uint32_t *__restrict__ sectors,
                    uint32_t *__restrict__ lines) {
    const uint32_t *row = trace + (size_t)blockIdx.x * words;
    uint32_t ns = 0, nl = 0;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) {
        const uint32_t b = row[i];
        ns += __popc(b);
        nl += __popc((b | (b >> 1) | (b >> 2) | (b >> 3)) & 0x11111111u);
    }
    __shared__ uint32_t s_ns, s_nl;
    if (threadIdx.x == 0) s_ns = s_nl = 0;
    __syncthreads();
    ns = __reduce_add_sync(0xffffffffu, ns);
    nl = __reduce_add_sync(0xffffffffu, nl);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_ns, ns); atomicAdd(&s_nl, nl); }
    __syncthreads();
    if (threadIdx.x == 0) { sectors[blockIdx.x] = s_ns; lines[blockIdx.x] = s_nl; }
}

static void launch_dup_seq(wd_ctx *ctx, bool all_bcl, const TileDesc *descs, const int32_t *rows, unsigned long long n_rows,
                           const uint32_t *slot_well, const uint32_t *tgt_off, int len, uint8_t *codes) {
    const unsigned long long *g_off = ctx->order_dev.as<unsigned long long>();
    const uint8_t *g_kind = ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8;
    const unsigned blocks = (unsigned)((2 * n_rows * (unsigned long long)len + 127) / 128);
    if (all_bcl) dup_seq_kernel<true><<<blocks, 128, 0, ctx->stream>>>(descs, rows, n_rows, slot_well, tgt_off, g_off, g_kind, len, codes);
    else dup_seq_kernel<false><<<blocks, 128, 0, ctx->stream>>>(descs, rows, n_rows, slot_well, tgt_off, g_off, g_kind, len, codes);
    ctx->launches++;
}

static void launch_trace_reduce(wd_ctx *ctx, const uint32_t *trace, uint32_t words, uint32_t n_rows, uint32_t *sectors,
                                uint32_t *lines) {
    trace_reduce_kernel<<<n_rows, 256, 0, ctx->stream>>>(trace, words, sectors, lines);
    ctx->launches++;
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
    const void *before = ctx->order_dev.p;
    WD_TRY(ctx->order_dev.reserve((size_t)MAX_ORDER * 9));
    // a lane is counted with the same plane order call after call: upload it when it changes only (small
    // pageable copies in front of a 0.4 ms kernel are not free)
    if (ctx->order_dev.p != before || off != ctx->order_off || kind != ctx->order_kind) {
        WD_CUDA(cudaMemcpyAsync(ctx->order_dev.p, off.data(), (size_t)len * 8, cudaMemcpyHostToDevice, st));
        WD_CUDA(cudaMemcpyAsync(ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8, kind.data(), (size_t)len,
                                cudaMemcpyHostToDevice, st));
        WD_CUDA(cudaStreamSynchronize(st));          // the sources are pageable and local
        ctx->order_off.swap(off);
        ctx->order_kind.swap(kind);
    }
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
    const void *before = ctx->descs.p;
    WD_TRY(ctx->descs.reserve((size_t)n_tiles * sizeof(TileDesc)));
    const size_t bytes = (size_t)n_tiles * sizeof(TileDesc);
    if (ctx->descs.p != before || ctx->descs_host.size() != bytes || memcmp(ctx->descs_host.data(), h.data(), bytes) != 0) {
        WD_CUDA(cudaMemcpyAsync(ctx->descs.p, h.data(), bytes, cudaMemcpyHostToDevice, st));
        WD_CUDA(cudaStreamSynchronize(st));          // `h` is pageable and local
        ctx->descs_host.assign(reinterpret_cast<const uint8_t *>(h.data()), reinterpret_cast<const uint8_t *>(h.data()) + bytes);
    }
    return WD_OK;
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
    switch (words) {
        case 1: launch_unpack_w<1>(ctx, ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 2: launch_unpack_w<2>(ctx, ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 4: launch_unpack_w<4>(ctx, ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        case 8: launch_unpack_w<8>(ctx, ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
        default: launch_unpack_w<16>(ctx, ctx->gs_packed.as<uint64_t>(), n_idx, seq_len, d_codes, d_pf); break;
    }
    WD_CUDA(cudaGetLastError());
    WD_CUDA(cudaMemcpyAsync(codes, d_codes, (size_t)n_idx * seq_len, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaMemcpyAsync(pf, d_pf, (size_t)n_idx, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    return WD_OK;
}

static void launch_count_any(wd_ctx *ctx, int words, const CountArgs &a, int n_tiles, int mode, bool all_bcl) {
    switch (words) {
        case 1: launch_count_w<1>(ctx, a, n_tiles, mode, all_bcl); break;
        case 2: launch_count_w<2>(ctx, a, n_tiles, mode, all_bcl); break;
        case 4: launch_count_w<4>(ctx, a, n_tiles, mode, all_bcl); break;
        case 8: launch_count_w<8>(ctx, a, n_tiles, mode, all_bcl); break;
        default: launch_count_w<16>(ctx, a, n_tiles, mode, all_bcl); break;
    }
}

// Head planes of host-mapped tiles k0 .. k1-1 -> ctx->head by DMA on the copy stream.  The staging pipeline
// lays a batch out as [tile][plane][stride] in one page-locked block, so the head planes of a run of tiles are
// `n_head * stride` contiguous bytes every `pitch` bytes: ONE strided copy per run instead of one per tile and
// plane (eight GPUs issuing thousands of small copies through one host is what SCALE_r01 showed).
static int copy_head_planes(wd_ctx *ctx, int first_slot, int k0, int k1, const int32_t *order, int n_head, size_t hstride) {
    bool consecutive = true;
    for (int j = 1; j < n_head; ++j) consecutive = consecutive && order[j] == order[0] + j;
    uint8_t *head = ctx->head.as<uint8_t>();
    int k = k0;
    while (k < k1) {
        const TileSlot &s = ctx->slots[first_slot + k];
        if (!consecutive || s.stride != hstride) {
            for (int j = 0; j < n_head; ++j) {
                const size_t bytes = std::min(s.stride, hstride);
                WD_CUDA(cudaMemcpyAsync(head + ((size_t)k * n_head + j) * hstride, s.mapped_host + (size_t)order[j] * s.stride, bytes,
                                        cudaMemcpyHostToDevice, ctx->copy_stream));
                ctx->last_h2d_bytes += bytes;
            }
            ++k;
            continue;
        }
        // longest run of tiles at a constant pitch
        int run = 1;
        ptrdiff_t pitch = 0;
        if (k + 1 < k1) pitch = ctx->slots[first_slot + k + 1].mapped_host - s.mapped_host;
        if (pitch >= (ptrdiff_t)(n_head * hstride) && pitch < (ptrdiff_t)0x7fffffff) {
            while (k + run < k1 && ctx->slots[first_slot + k + run].mapped_host - ctx->slots[first_slot + k + run - 1].mapped_host == pitch &&
                   ctx->slots[first_slot + k + run].stride == hstride)
                ++run;
        }
        const uint8_t *src = s.mapped_host + (size_t)order[0] * s.stride;
        uint8_t *dst = head + (size_t)k * n_head * hstride;
        if (run == 1) WD_CUDA(cudaMemcpyAsync(dst, src, (size_t)n_head * hstride, cudaMemcpyHostToDevice, ctx->copy_stream));
        else WD_CUDA(cudaMemcpy2DAsync(dst, (size_t)n_head * hstride, src, (size_t)pitch, (size_t)n_head * hstride, (size_t)run,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->last_h2d_bytes += (uint64_t)run * n_head * hstride;
        k += run;
    }
    return WD_OK;
}

// mode: WD_MODE_FUSED, WD_MODE_TWO_PASS (logs), WD_MODE_FUSED_LOG; trace != null: measurement build of the fused kernel
static int count_run(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e, int hamming,
                     int mode, int want_per_target, uint32_t *trace, uint32_t trace_words) {
    TargetList &tl = ctx->targets;
    if (tl.t == 0) WD_FAIL(WD_E_ARG, "wd_count: no target list loaded");
    if (mode != WD_MODE_FUSED && mode != WD_MODE_TWO_PASS && mode != WD_MODE_FUSED_LOG)
        WD_FAIL(WD_E_ARG, "wd_count: mode must be 0 (fused), 1 (two-pass) or 2 (fused, duplicate pairs logged)");
    if (n_tiles > 65535) WD_FAIL(WD_E_ARG, "wd_count: at most 65535 tiles per call");
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, first_slot, n_tiles, order, seq_len, &all_bcl, &any_excl));
    for (int k = 0; k < n_tiles; ++k) {
        const TileSlot &s = ctx->slots[first_slot + k];
        if (tl.max_well >= s.n)
            WD_FAIL(WD_E_INDEX, "Requested cluster %u is out of range.  Highest on this tile is %u.", tl.max_well, s.n - 1);
    }
    if (any_excl) {
        std::vector<int> ids(n_tiles);
        for (int k = 0; k < n_tiles; ++k) ids[k] = first_slot + k;
        WD_TRY(filter_rank(ctx, ids.data(), n_tiles));
        for (int k = 0; k < n_tiles; ++k) {
            const int id = first_slot + k;
            // every excluded block of a tile holds its PF wells (cbcl_read.py:130-131): the kernels compare the
            // count of the first with the filter's total, so the others have to agree with the first
            const TileSlot &s = ctx->slots[id];
            int first = -1;
            for (int p = 0; p < s.n_planes; ++p) {
                if (s.kind[p] != WD_PLANE_CBCL_EXCL) continue;
                if (first < 0) first = p;
                else if (s.n_block[p] != s.n_block[first])
                    WD_FAIL(WD_E_ASSERT, "tile slot %d: excluded CBCL blocks of planes %d and %d hold %u and %u clusters", id, first, p,
                            s.n_block[first], s.n_block[p]);
            }
        }
    }
    cudaStream_t st = ctx->stream;
    const int L = tl.levels;
    const size_t width = 1 + 5 * (size_t)L;
    ctx->dup_raw_valid = false;
    const bool fused = mode != WD_MODE_TWO_PASS;
    const bool logged = mode != WD_MODE_FUSED;
    const Tuning &tu = ctx->tuning;
    const bool over_pcie = ctx->slots[first_slot].mapped != nullptr;
    for (int k = 0; k < n_tiles; ++k)
        if ((ctx->slots[first_slot + k].mapped != nullptr) != over_pcie)
            WD_FAIL(WD_E_ARG, "wd_count: host-mapped and staged tile slots cannot share one call");

    // host-mapped tiles: the planes of the first compared positions go to HBM by DMA (below)
    int n_head = 0, n_groups = 1;
    size_t hstride = 0;
    if (fused && over_pcie) {
        // How many leading planes go by DMA: one.  On a single GPU one, two or three planes give the same step
        // (31.5 / 32.0 / 33.0 ms: copies and sector pulls share the PCIe read path and overlap only in part), on
        // eight GPUs behind one host every further plane costs 12.5 ms (49 / 61 / 74 / 87 ms for 1 / 2 / 3 / 4
        // planes) -- profiles/r02_notes.md.  The rate of the copies is measured for the record.
        if (ctx->dma_pending && cudaEventQuery(ctx->dma_ev1) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->dma_ev0, ctx->dma_ev1) == cudaSuccess && ms > 0.f)
                ctx->dma_gbps = (double)ctx->dma_bytes_timed / ms / 1e6;
            ctx->dma_pending = false;
        }
        cudaGetLastError();
        n_head = tu.head_planes >= 0 ? tu.head_planes : 1;
        n_head = std::max(0, std::min(n_head, std::min(seq_len, 8)));
        n_groups = std::max(1, std::min(tu.head_groups > 0 ? tu.head_groups : 16, n_tiles));
        hstride = ctx->slots[first_slot].stride;          // prepare_order: the same for every tile of the batch
    }
    if (n_head > 0) {
        if ((size_t)n_tiles * n_head * hstride > ctx->head.cap) WD_CUDA(cudaStreamSynchronize(st));
        WD_TRY(ctx->head.reserve((size_t)n_tiles * n_head * hstride));
    }
    for (int k = 0; k < n_tiles; ++k) {
        TileSlot &s = ctx->slots[first_slot + k];
        s.head_ptr = n_head > 0 ? ctx->head.as<uint8_t>() + (size_t)k * n_head * hstride : nullptr;
        s.head_stride = hstride;
    }
    WD_TRY(upload_descs(ctx, first_slot, n_tiles));

    // counter rows + two status words behind them
    WD_TRY(ctx->counters.reserve(((size_t)n_tiles * width + 2) * 8));
    WD_CUDA(cudaMemsetAsync(ctx->counters.p, 0, ((size_t)n_tiles * width + 2) * 8, st));
    if (want_per_target) WD_TRY(ctx->per_target.reserve((size_t)n_tiles * tl.t * (1 + 2 * L) * 4));
    const int words = words_for(seq_len);
    if (trace != nullptr && (words != 1 || L > 5))
        WD_FAIL(WD_E_ARG, "wd_count_trace_sectors: the measurement build covers at most 64 compared symbols and 5 levels");

    CountArgs a;
    memset(&a, 0, sizeof(a));
    a.descs = ctx->descs.as<TileDesc>();
    a.tgt_off = tl.tgt_off.as<uint32_t>();
    a.slot_well = tl.slot_well.as<uint32_t>();
    a.slot_csr = tl.slot_csr.as<uint32_t>();
    a.level_len = tl.level_len.as<uint32_t>();
    a.visit = tu.visit_order == 0 ? nullptr : tl.visit.as<uint32_t>();
    a.slot_level = tl.slot_level.as<uint8_t>();
    a.g_off = ctx->order_dev.as<unsigned long long>();
    a.g_kind = ctx->order_dev.as<uint8_t>() + (size_t)MAX_ORDER * 8;
    a.per_target = want_per_target ? ctx->per_target.as<int32_t>() : nullptr;
    a.counters = ctx->counters.as<unsigned long long>();
    a.status = ctx->counters.as<unsigned long long>() + (size_t)n_tiles * width;
    a.trace = trace;
    a.trace_words = trace_words;
    a.t = tl.t;
    a.n_slots = tl.n_slots;
    a.levels = L;
    a.len = seq_len;
    a.e = e;
    a.hamming = hamming;
    // early-exit schedule of the fused kernel (cycles read per round: first, later), from the sweeps in
    // profiles/: planes in HBM are bound by instruction issue and latency -> few long rounds; planes pulled
    // across PCIe (wd_tile_map_host) are bound by the number of sector requests -> many short rounds
    a.step0 = over_pcie ? (n_head > 0 ? n_head : 4) : 8;       // host-mapped: the first round entirely from HBM
    a.step1 = over_pcie ? 1 : 4;
    a.cchunk = 0;                // the centre is read exactly as far as each round looks ahead
    if (tu.step0 >= 1 && tu.step0 <= 8) a.step0 = tu.step0;
    if (tu.step1 >= 1 && tu.step1 <= 8) a.step1 = tu.step1;
    if (tu.centre_chunk == 8 || tu.centre_chunk == 16 || tu.centre_chunk == 32) a.cchunk = tu.centre_chunk;
    a.n_head = n_head;
    // Targets per CTA.  A CTA's warps share its targets and wait for the slowest at the end; CTAs come in waves of
    // sm_count x 8.  Planes in HBM (sweeps of 16..128 in profiles/r02_notes.md section 2): a 96-tile lane is fastest
    // with 32 (6.4 waves; 64 = 3.2 waves is 4 % slower), the 704-tile CBCL lane with 128 (10 % faster than 32): as many
    // as leave six waves, between 32 and 128.  Host-mapped tiles: the sector pulls across PCIe are bound by the
    // requests in flight, and fewer, longer-lived CTAs are faster (BCL lane: 31.0 ms with 256 or 128, 32.2 with 64, 34.7 with 32;
    // 704-tile CBCL lane: 251 ms with 256, 286 with 128, 289 with 64)
    {
        const size_t resident = (size_t)ctx->sm_count * 8;
        const size_t per_launch = (size_t)tl.t * (size_t)std::max(1, n_tiles / std::max(1, n_groups));
        size_t tpb = over_pcie ? 256 : std::min<size_t>(128, per_launch / (6 * resident));
        tpb = std::max<size_t>(32, tpb / 8 * 8);
        a.tpb = (int)tpb;
    }
    if (tu.targets_per_cta >= 8) a.tpb = tu.targets_per_cta;
    // wd_set_tuning only: cap the CTAs resident per SM by reserving shared memory (227 KB per SM, ~11 KB per CTA static)
    a.pad_smem = tu.ctas_per_sm > 0 ? std::min(200 * 1024, std::max(0, (227 * 1024) / tu.ctas_per_sm - 12 * 1024) / 1024 * 1024) : 0;

    if (!fused) {
        WD_TRY(ctx->packed.reserve((size_t)n_tiles * tl.n_slots * words * PACK_STRIDE * 8));
        launch_gather_any(ctx, words, all_bcl, a.descs, a.slot_well, tl.n_slots, n_tiles, seq_len, ctx->packed.as<uint64_t>());
        a.packed = ctx->packed.as<uint64_t>();
    }
    if (logged) {
        // The log holds what a run of real data produces many times over; a run that produces more (wd_dup_pairs
        // learns the count) makes wd_dup_pairs grow the buffer and repeat the count -- every pair is logged, as in
        // the reference, however many there are.
        const size_t cap = std::max(ctx->dup_cap_wanted, (size_t)n_tiles * tl.n_slots / 8 + 65536);
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
    ctx->last_h2d_bytes = 0;
    if (n_head > 0) {
        // Host-mapped tiles: the planes every ring well is read from -- the first positions -- go to HBM
        // by DMA (bandwidth-bound, ~52 GB/s) while the kernel pulls only what the survivors need of the
        // later planes as 32-byte sector reads (request-bound, ~0.4 G requests/s): the two limits of the
        // PCIe path are used side by side, group of tiles after group of tiles.
        while ((int)ctx->copy_events.size() < n_groups + 1) {
            cudaEvent_t ev;
            WD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->copy_events.push_back(ev);
        }
        // the copies may not overtake earlier work on the compute stream that still reads the head buffer
        WD_CUDA(cudaEventRecord(ctx->copy_events[n_groups], st));
        WD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[n_groups], 0));
        if (ctx->dma_ev0 == nullptr) {
            WD_CUDA(cudaEventCreate(&ctx->dma_ev0));
            WD_CUDA(cudaEventCreate(&ctx->dma_ev1));
        }
        const bool time_dma = !ctx->dma_pending && trace == nullptr;
        if (time_dma) WD_CUDA(cudaEventRecord(ctx->dma_ev0, ctx->copy_stream));
        const size_t row = 1 + 2 * (size_t)L;
        for (int g = 0; g < n_groups; ++g) {
            const int t0 = (int)((long long)n_tiles * g / n_groups), t1 = (int)((long long)n_tiles * (g + 1) / n_groups);
            WD_TRY(copy_head_planes(ctx, first_slot, t0, t1, order, n_head, hstride));
            WD_CUDA(cudaEventRecord(ctx->copy_events[g], ctx->copy_stream));
            WD_CUDA(cudaStreamWaitEvent(st, ctx->copy_events[g], 0));
            CountArgs ag = a;
            ag.descs = a.descs + t0;
            ag.counters = a.counters + (size_t)t0 * width;
            ag.tile_base = (uint32_t)t0;
            if (a.per_target) ag.per_target = a.per_target + (size_t)t0 * tl.t * row;
            if (a.trace) ag.trace = a.trace + (size_t)t0 * seq_len * trace_words;
            launch_count_any(ctx, words, ag, t1 - t0, 0, all_bcl);
        }
        if (time_dma) {
            WD_CUDA(cudaEventRecord(ctx->dma_ev1, ctx->copy_stream));
            ctx->dma_pending = true;
            ctx->dma_bytes_timed = ctx->last_h2d_bytes;
        }
    } else {
        launch_count_any(ctx, words, a, n_tiles, fused ? 0 : 1, all_bcl);
    }
    WD_CUDA(cudaGetLastError());
    ctx->last_tiles = n_tiles;
    ctx->last_levels = L;
    ctx->last_t = tl.t;
    ctx->last_first_slot = first_slot;
    ctx->last_per_target = want_per_target != 0;
    ctx->last_all_bcl = all_bcl;
    ctx->last_seq_len = seq_len;
    ctx->last_n_head = n_head;
    return WD_OK;
}

int count_async(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e, int hamming,
                int mode, int want_per_target) {
    ctx->dup_cap_wanted = 0;
    WD_TRY(count_run(ctx, first_slot, n_tiles, order, seq_len, e, hamming, mode, want_per_target, nullptr, 0));
    if (order != ctx->last_order.data()) ctx->last_order.assign(order, order + seq_len);
    ctx->last_e = e;
    ctx->last_hamming = hamming;
    ctx->last_mode = mode;
    return WD_OK;
}

// Raw rows (tile, target, slot, distance) of the last logged count, in no particular order.  If there were more
// pairs than the log had room for, the log is grown to the size the kernels reported and the count is repeated
// (the tile slots still hold their planes): the reference logs every pair, however many.
int dup_rows_fetch(wd_ctx *ctx, std::vector<int32_t> &raw) {
    cudaStream_t st = ctx->stream;
    if (ctx->dup_raw_valid) {
        raw = ctx->dup_raw;
        return WD_OK;
    }
    for (int attempt = 0; attempt < 3; ++attempt) {
        unsigned long long n = 0;
        WD_CUDA(cudaMemcpyAsync(&n, ctx->dup_count.p, 8, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaStreamSynchronize(st));
        if (n <= ctx->dup_cap) {
            raw.resize((size_t)n * 4);
            if (n) WD_CUDA(cudaMemcpyAsync(raw.data(), ctx->dup_rows.p, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
            WD_CUDA(cudaStreamSynchronize(st));
            ctx->dup_raw = raw;
            ctx->dup_raw_valid = true;
            return WD_OK;
        }
        ctx->dup_cap_wanted = (size_t)n + (size_t)n / 16 + 1024;
        WD_TRY(count_run(ctx, ctx->last_first_slot, ctx->last_tiles, ctx->last_order.data(), ctx->last_seq_len, ctx->last_e,
                         ctx->last_hamming, ctx->last_mode, ctx->last_per_target ? 1 : 0, nullptr, 0));
    }
    WD_FAIL(WD_E_CAPACITY, "duplicate-pair log kept overflowing");
}

// codes[row][centre, well][seq_len] for raw rows (same order)
int dup_seqs(wd_ctx *ctx, const std::vector<int32_t> &raw, std::vector<uint8_t> &codes) {
    const size_t n = raw.size() / 4;
    const int len = ctx->last_seq_len;
    codes.resize(n * 2 * (size_t)len);
    if (n == 0) return WD_OK;
    cudaStream_t st = ctx->stream;
    WD_TRY(ctx->dup_codes.reserve(n * 2 * (size_t)len));
    // the plane order and the tile descriptors of that count, in case a wd_get_seqs came in between (both are
    // uploaded only if they differ from what the device holds)
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, ctx->last_first_slot, ctx->last_tiles, ctx->last_order.data(), len, &all_bcl, &any_excl));
    WD_TRY(upload_descs(ctx, ctx->last_first_slot, ctx->last_tiles));
    // the rows are still in the log buffer, in the order they were fetched in
    launch_dup_seq(ctx, ctx->last_all_bcl, ctx->descs.as<TileDesc>(), ctx->dup_rows.as<int32_t>(), n,
                   ctx->targets.slot_well.as<uint32_t>(), ctx->targets.tgt_off.as<uint32_t>(), len, ctx->dup_codes.as<uint8_t>());
    WD_CUDA(cudaGetLastError());
    WD_CUDA(cudaMemcpyAsync(codes.data(), ctx->dup_codes.p, codes.size(), cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    return WD_OK;
}

// Measurement hook: the fused kernel's reads, recorded.  sectors / lines [n_tiles][seq_len] = distinct 32-byte
// sectors / 128-byte lines of the plane of each compared position that the kernel asks for, with the schedule
// and early exits wd_count would use on resident planes.
int count_trace(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e, int hamming,
                uint32_t *sectors, uint32_t *lines) {
    if (first_slot < 0 || n_tiles < 1 || (size_t)first_slot + n_tiles > ctx->slots.size())
        WD_FAIL(WD_E_ARG, "tile slots %d..%d have not been begun", first_slot, first_slot + n_tiles - 1);
    cudaStream_t st = ctx->stream;
    size_t max_stride = 0;
    for (int k = 0; k < n_tiles; ++k) max_stride = std::max(max_stride, ctx->slots[first_slot + k].stride);
    const uint32_t words = (uint32_t)((max_stride / 32 + 31) / 32);
    const size_t per_tile = (size_t)seq_len * words * 4;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_tiles, ((size_t)1 << 30) / per_tile));
    WD_CUDA(cudaStreamSynchronize(st));
    WD_TRY(ctx->trace.reserve((size_t)chunk * per_tile));
    WD_TRY(ctx->trace_counts.reserve((size_t)chunk * seq_len * 8));
    for (int k0 = 0; k0 < n_tiles; k0 += chunk) {
        const int nk = std::min(chunk, n_tiles - k0);
        WD_CUDA(cudaMemsetAsync(ctx->trace.p, 0, (size_t)nk * per_tile, st));
        ctx->dup_cap_wanted = 0;
        WD_TRY(count_run(ctx, first_slot + k0, nk, order, seq_len, e, hamming, WD_MODE_FUSED, 0, ctx->trace.as<uint32_t>(), words));
        uint32_t *d_sec = ctx->trace_counts.as<uint32_t>(), *d_lin = d_sec + (size_t)nk * seq_len;
        launch_trace_reduce(ctx, ctx->trace.as<uint32_t>(), words, (uint32_t)(nk * seq_len), d_sec, d_lin);
        WD_CUDA(cudaGetLastError());
        WD_CUDA(cudaMemcpyAsync(sectors + (size_t)k0 * seq_len, d_sec, (size_t)nk * seq_len * 4, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaMemcpyAsync(lines + (size_t)k0 * seq_len, d_lin, (size_t)nk * seq_len * 4, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaStreamSynchronize(st));
    }
    ctx->last_tiles = 0;          // the counters of a traced run are not a result to fetch
    return WD_OK;
}

// Two buffers alternate, so that the all-reduce of one step (on the communication stream, wd_comm.cc) runs
// under the counting kernels of the next; `keep` adds to the buffer of the previous call instead (a rank that
// counts its tiles in several batches publishes each batch into the same rows buffer).
int publish_counters(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles, int n_rows_total, bool keep) {
    // n_tiles == 0: a rank without tiles still brings a zeroed array to the all-reduce
    if (n_tiles != 0 && n_tiles != ctx->last_tiles)
        WD_FAIL(WD_E_ARG, "wd_publish_counters: last wd_count covered %d tiles, not %d", ctx->last_tiles, n_tiles);
    if (ctx->targets.t == 0) WD_FAIL(WD_E_ARG, "wd_publish_counters: no target list loaded");
    const int width = 1 + 5 * ctx->targets.levels;
    for (int i = 0; i < n_tiles; ++i)
        if (tile_row[i] >= n_rows_total || lane_row[i] >= n_rows_total)
            WD_FAIL(WD_E_ARG, "wd_publish_counters: row index out of range");
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)n_rows_total * width;
    if (keep && (n != ctx->publish_n || ctx->publish.p == nullptr))
        WD_FAIL(WD_E_ARG, "wd_publish_add: no buffer of %d rows has been published yet", n_rows_total);
    if (ctx->pub_ready == nullptr) {
        WD_CUDA(cudaEventCreateWithFlags(&ctx->pub_ready, cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) WD_CUDA(cudaEventCreateWithFlags(&ctx->comm_done[b], cudaEventDisableTiming));
    }
    const size_t need = 2 * n * 8 + (size_t)n_tiles * 8;
    if (need > ctx->publish.cap || n != ctx->publish_n) {
        // growing frees the old block, another row count moves the second buffer: nothing may still be reading or reducing them
        WD_CUDA(cudaStreamSynchronize(st));
        if (ctx->comm_stream) WD_CUDA(cudaStreamSynchronize(ctx->comm_stream));
        ctx->comm_pending[0] = ctx->comm_pending[1] = false;
        ctx->publish_map.clear();
    }
    WD_TRY(ctx->publish.reserve(need));
    if (!keep) ctx->publish_cur ^= 1;
    const int cur = ctx->publish_cur;
    unsigned long long *buf = ctx->publish.as<unsigned long long>() + (size_t)cur * n;
    int32_t *rows = reinterpret_cast<int32_t *>(ctx->publish.as<unsigned long long>() + 2 * n);
    if (ctx->comm_pending[cur]) {
        WD_CUDA(cudaStreamWaitEvent(st, ctx->comm_done[cur], 0));      // its previous all-reduce has to be over
        ctx->comm_pending[cur] = false;
    }
    if (!keep) WD_CUDA(cudaMemsetAsync(buf, 0, n * 8, st));
    // the row maps are the same step after step (a rank keeps its tiles): upload them when they change only --
    // two small pageable copies cost more than the kernel they feed
    std::vector<int32_t> map(tile_row, tile_row + n_tiles);
    map.insert(map.end(), lane_row, lane_row + n_tiles);
    if (!map.empty() && (n != ctx->publish_n || map != ctx->publish_map)) {
        WD_CUDA(cudaMemcpyAsync(rows, map.data(), map.size() * 4, cudaMemcpyHostToDevice, st));
        WD_CUDA(cudaStreamSynchronize(st));          // `map` is pageable and local
        ctx->publish_map.swap(map);
    }
    const int total = n_tiles * width;
    if (total > 0) {
        publish_kernel<<<(total + 255) / 256, 256, 0, st>>>(ctx->counters.as<unsigned long long>(), rows, rows + n_tiles,
                                                             n_tiles, width, buf);
        ctx->launches++;
        WD_CUDA(cudaGetLastError());
    }
    ctx->publish_n = n;
    return WD_OK;
}

// dense pass of exhaustive mode (wd_exhaustive.cu): every well of one tile packed once, in index
// order (coalesced plane reads), into ctx->x_packed
int pack_dense(wd_ctx *ctx, int slot, const int32_t *order, int seq_len, int *words_out) {
    TileSlot &s = ctx->slots[slot];
    bool all_bcl, any_excl;
    WD_TRY(prepare_order(ctx, slot, 1, order, seq_len, &all_bcl, &any_excl));
    if (any_excl) WD_TRY(filter_rank(ctx, &slot, 1));
    WD_TRY(upload_descs(ctx, slot, 1));
    const int words = words_for(seq_len);
    WD_TRY(ctx->x_packed.reserve((size_t)s.n * words * PACK_STRIDE * 8));
    launch_gather_any(ctx, words, all_bcl, ctx->descs.as<TileDesc>(), nullptr, s.n, 1, seq_len, ctx->x_packed.as<uint64_t>());
    WD_CUDA(cudaGetLastError());
    *words_out = words;
    return WD_OK;
}

}  // namespace wd
