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
    memset(&d, 0, sizeof(d));                     // descriptors are compared bytewise (upload_descs)
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
    a.step1 = over_pcie ? 1 : 4;
    a.cchunk = over_pcie ? 8 : 16;
    if (const char *cc = getenv("WELLDUP_CENTRE_CHUNK")) {
        const int v = atoi(cc);
        if (v == 8 || v == 16 || v == 32) a.cchunk = v;
    }
    if (const char *sch = getenv("WELLDUP_STEPS")) {
        int s0 = 0, s1 = 0;
        if (sscanf(sch, "%d,%d", &s0, &s1) == 2 && s0 >= 1 && s0 <= 8 && s1 >= 1 && s1 <= 8) {
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
        n_head = std::max(0, std::min(n_head, std::min(seq_len, 8)));
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
    const void *before = ctx->publish.p;
    WD_TRY(ctx->publish.reserve(n * 8 + (size_t)n_tiles * 8));
    int32_t *rows = reinterpret_cast<int32_t *>(ctx->publish.as<unsigned long long>() + n);
    WD_CUDA(cudaMemsetAsync(ctx->publish.p, 0, n * 8, st));
    // the row maps are the same step after step (a rank keeps its tiles): upload them when they change only --
    // two small pageable copies cost more than the kernel they feed
    std::vector<int32_t> map(tile_row, tile_row + n_tiles);
    map.insert(map.end(), lane_row, lane_row + n_tiles);
    if (ctx->publish.p != before || n != ctx->publish_n || map != ctx->publish_map) {
        WD_CUDA(cudaMemcpyAsync(rows, map.data(), map.size() * 4, cudaMemcpyHostToDevice, st));
        WD_CUDA(cudaStreamSynchronize(st));          // `map` is pageable and local
        ctx->publish_map.swap(map);
    }
    const int total = n_tiles * width;
    publish_kernel<<<(total + 255) / 256, 256, 0, st>>>(ctx->counters.as<unsigned long long>(), rows, rows + n_tiles,
                                                         n_tiles, width, ctx->publish.as<unsigned long long>());
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
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
