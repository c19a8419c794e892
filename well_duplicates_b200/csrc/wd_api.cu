// extern "C" surface of libwelldup.so (see include/welldup.h): context, tile
// staging, target list, result fetch.  Kernels live in wd_stage1.cu and
// wd_stage23.cu.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "wd_common.cuh"

namespace wd {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p != nullptr) return WD_OK;
    if (bytes == 0) bytes = 16;
    if (p) {
        cudaError_t e = cudaFree(p);   // implicit device sync: no kernel still reads the old block
        p = nullptr;
        cap = 0;
        if (e != cudaSuccess) WD_FAIL(WD_E_CUDA, "cudaFree failed: %s", cudaGetErrorString(e));
    }
    const size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        WD_FAIL(WD_E_CUDA, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return WD_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static int check_slot(wd_ctx *ctx, int slot, const char *who) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "%s: null context", who);
    if (slot < 0 || (size_t)slot >= ctx->slots.size() || ctx->slots[slot].n == 0)
        WD_FAIL(WD_E_ARG, "%s: tile slot %d has not been begun", who, slot);
    return WD_OK;
}

}  // namespace wd

using namespace wd;

extern "C" {

int wd_abi_version(void) { return WD_ABI_VERSION; }

const char *wd_last_error(void) { return g_err; }

int wd_create(int device, wd_ctx **out) {
    if (out == nullptr) WD_FAIL(WD_E_ARG, "wd_create: null output pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        WD_FAIL(WD_E_CUDA, "wd_create: no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) WD_FAIL(WD_E_ARG, "wd_create: device %d not in 0..%d", device, n_dev - 1);
    WD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    WD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        WD_FAIL(WD_E_CUDA, "wd_create: device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major,
                prop.minor);
    wd_ctx *ctx = new wd_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        WD_FAIL(WD_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        WD_FAIL(WD_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    // scattered 1-byte gathers: ask L2 not to pull whole lines from HBM (a hint this part ignores -- the plane
    // loads carry their own fill size, wd_kernels23.cuh: ld_plane_u8; failure is harmless)
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaGetLastError();
    *out = ctx;
    return WD_OK;
}

int wd_set_l2_fetch_granularity(wd_ctx *ctx, int bytes, int *previous) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_set_l2_fetch_granularity: null context");
    if (bytes != 32 && bytes != 64 && bytes != 128) WD_FAIL(WD_E_ARG, "L2 fetch granularity must be 32, 64 or 128");
    WD_CUDA(cudaSetDevice(ctx->device));
    size_t old = 0;
    WD_CUDA(cudaDeviceGetLimit(&old, cudaLimitMaxL2FetchGranularity));
    if (previous) *previous = (int)old;
    WD_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return WD_OK;
}

int wd_set_tuning(wd_ctx *ctx, const wd_tuning *t) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_set_tuning: null context");
    Tuning tu;
    if (t != nullptr) {
        if (t->step0 < 0 || t->step0 > 8 || t->step1 < 0 || t->step1 > 8) WD_FAIL(WD_E_ARG, "wd_set_tuning: rounds read 1..8 cycles (0 = default)");
        if (t->centre_chunk != 0 && t->centre_chunk != 8 && t->centre_chunk != 16 && t->centre_chunk != 32)
            WD_FAIL(WD_E_ARG, "wd_set_tuning: centre_chunk is 8, 16 or 32 (0 = default)");
        if (t->head_planes > 8) WD_FAIL(WD_E_ARG, "wd_set_tuning: at most 8 head planes");
        tu.step0 = t->step0;
        tu.step1 = t->step1;
        tu.centre_chunk = t->centre_chunk;
        tu.head_planes = t->head_planes;
        tu.head_groups = t->head_groups;
        tu.visit_order = t->visit_order;
        if (t->targets_per_cta != 0 && (t->targets_per_cta < 8 || t->targets_per_cta > 256 || t->targets_per_cta % 8 != 0))
            WD_FAIL(WD_E_ARG, "wd_set_tuning: targets_per_cta is a multiple of 8 in 8..256 (0 = default)");
        if (t->ctas_per_sm < 0 || t->ctas_per_sm > 8) WD_FAIL(WD_E_ARG, "wd_set_tuning: ctas_per_sm is 1..8 (0 = no limit)");
        tu.targets_per_cta = t->targets_per_cta;
        tu.ctas_per_sm = t->ctas_per_sm;
    }
    ctx->tuning = tu;
    return WD_OK;
}

int wd_destroy(wd_ctx *ctx) {
    if (ctx == nullptr) return WD_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    comm_destroy(ctx);
    DevBuf *bufs[] = {&ctx->xy, &ctx->px, &ctx->py, &ctx->bbox, &ctx->cell_start, &ctx->cell_cursor, &ctx->cell_wells,
                      &ctx->scan_tmp, &ctx->q_centres, &ctx->q_counts, &ctx->q_offsets, &ctx->q_idx, &ctx->q_tmp,
                      &ctx->q_flag, &ctx->descs, &ctx->order_dev, &ctx->packed, &ctx->per_target, &ctx->counters,
                      &ctx->publish, &ctx->dup_rows, &ctx->dup_count, &ctx->gs_idx, &ctx->gs_packed, &ctx->gs_codes,
                      &ctx->targets.tgt_off, &ctx->targets.slot_well, &ctx->targets.slot_level, &ctx->targets.slot_csr,
                      &ctx->targets.level_len, &ctx->targets.visit, &ctx->head, &ctx->trace, &ctx->trace_counts, &ctx->dup_codes, &ctx->rank_jobs, &ctx->x_packed, &ctx->x_counts, &ctx->x_pre, &ctx->x_ringlen, &ctx->x_flags, &ctx->x_tally, &ctx->x_work};
    for (DevBuf *b : bufs) b->release();
    for (TileSlot &s : ctx->slots) {
        s.planes.release(); s.filter.release(); s.pfmask.release(); s.pfrank.release();
        s.kind_dev.release();
    }
    for (cudaEvent_t ev : ctx->copy_events) cudaEventDestroy(ev);
    if (ctx->dma_ev0) cudaEventDestroy(ctx->dma_ev0);
    if (ctx->dma_ev1) cudaEventDestroy(ctx->dma_ev1);
    if (ctx->pub_ready) cudaEventDestroy(ctx->pub_ready);
    for (cudaEvent_t ev : ctx->comm_done) if (ev) cudaEventDestroy(ev);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return WD_OK;
}

int wd_set_stream(wd_ctx *ctx, void *cuda_stream) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_set_stream: null context");
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return WD_OK;
}

int wd_sync(wd_ctx *ctx) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_sync: null context");
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaStreamSynchronize(ctx->stream));
    return WD_OK;
}

int wd_host_alloc(size_t bytes, void **out) {
    if (out == nullptr) WD_FAIL(WD_E_ARG, "wd_host_alloc: null output pointer");
    WD_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return WD_OK;
}

int wd_host_free(void *p) {
    if (p) WD_CUDA(cudaFreeHost(p));
    return WD_OK;
}

int wd_last_count_h2d_bytes(wd_ctx *ctx, uint64_t *out) {
    if (ctx == nullptr || out == nullptr) WD_FAIL(WD_E_ARG, "wd_last_count_h2d_bytes: null argument");
    *out = ctx->last_h2d_bytes;
    return WD_OK;
}

int wd_last_count_staging(wd_ctx *ctx, int *head_planes, double *dma_gb_per_s) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_last_count_staging: null context");
    if (head_planes) *head_planes = ctx->last_n_head;
    if (dma_gb_per_s) *dma_gb_per_s = ctx->dma_gbps;
    return WD_OK;
}

int wd_launch_count(wd_ctx *ctx, uint64_t *out) {
    if (ctx == nullptr || out == nullptr) WD_FAIL(WD_E_ARG, "wd_launch_count: null argument");
    *out = ctx->launches;
    return WD_OK;
}

// ---- stage 1 ------------------------------------------------------------------------
int wd_locs_load(wd_ctx *ctx, const float *xy, uint32_t n) {
    if (ctx == nullptr || xy == nullptr) WD_FAIL(WD_E_ARG, "wd_locs_load: null argument");
    WD_CUDA(cudaSetDevice(ctx->device));
    return locs_load(ctx, xy, n);
}

int wd_locs_pixels(wd_ctx *ctx, int32_t *x, int32_t *y) {
    if (ctx == nullptr || x == nullptr || y == nullptr) WD_FAIL(WD_E_ARG, "wd_locs_pixels: null argument");
    if (ctx->n_locs == 0) WD_FAIL(WD_E_ARG, "wd_locs_pixels: call wd_locs_load first");
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaMemcpyAsync(x, ctx->px.p, (size_t)ctx->n_locs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    WD_CUDA(cudaMemcpyAsync(y, ctx->py.p, (size_t)ctx->n_locs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    WD_CUDA(cudaStreamSynchronize(ctx->stream));
    return WD_OK;
}

int wd_ring_query(wd_ctx *ctx, const uint32_t *centres, uint32_t t, int levels, uint32_t window_lo,
                  uint32_t window_hi, uint32_t *level_offsets, uint32_t *idx, size_t idx_cap, uint64_t *n_idx,
                  uint32_t *first_empty) {
    if (ctx == nullptr || level_offsets == nullptr || (t && centres == nullptr))
        WD_FAIL(WD_E_ARG, "wd_ring_query: null argument");
    WD_CUDA(cudaSetDevice(ctx->device));
    return ring_query(ctx, centres, t, levels, window_lo, window_hi, level_offsets, idx, idx_cap, n_idx, first_empty);
}

// ---- target list -----------------------------------------------------------------------
int wd_targets_load(wd_ctx *ctx, const uint32_t *centres, const uint32_t *level_offsets, const uint32_t *idx,
                    uint32_t t, int levels) {
    if (ctx == nullptr || centres == nullptr || level_offsets == nullptr || idx == nullptr)
        WD_FAIL(WD_E_ARG, "wd_targets_load: null argument");
    if (t == 0) WD_FAIL(WD_E_ARG, "wd_targets_load: empty target list");
    if (levels < 1 || levels > WD_MAX_LEVELS)
        WD_FAIL(WD_E_ARG, "wd_targets_load: levels must be 1..%d", WD_MAX_LEVELS);
    WD_CUDA(cudaSetDevice(ctx->device));
    const size_t nseg = (size_t)t * levels;
    const uint64_t n_ring = level_offsets[nseg];
    if (level_offsets[0] != 0) WD_FAIL(WD_E_ARG, "wd_targets_load: level_offsets[0] must be 0");
    const uint64_t n_slots64 = (uint64_t)t + n_ring;
    if (n_slots64 >= (1ull << 32)) WD_FAIL(WD_E_ARG, "wd_targets_load: too many wells (%llu)", (unsigned long long)n_slots64);
    const uint32_t n_slots = (uint32_t)n_slots64;
    std::vector<uint32_t> tgt_off(t + 1), slot_well(n_slots), slot_csr(n_slots), level_len(nseg);
    std::vector<uint8_t> slot_level(n_slots);
    std::vector<std::pair<uint32_t, uint32_t>> tmp;   // (well, csr position)
    uint32_t max_well = 0;
    uint32_t pos = 0;
    bool has_empty = false;
    for (uint32_t i = 0; i < t; ++i) {
        tgt_off[i] = pos;
        slot_well[pos] = centres[i];
        slot_level[pos] = 0;
        slot_csr[pos] = UINT32_MAX;
        max_well = std::max(max_well, centres[i]);
        ++pos;
        const uint32_t a = level_offsets[(size_t)i * levels], b = level_offsets[(size_t)(i + 1) * levels];
        if (b < a) WD_FAIL(WD_E_ARG, "wd_targets_load: level_offsets must be non-decreasing");
        tmp.clear();
        for (int l = 0; l < levels; ++l) {
            const uint32_t s = level_offsets[(size_t)i * levels + l], e = level_offsets[(size_t)i * levels + l + 1];
            if (e < s) WD_FAIL(WD_E_ARG, "wd_targets_load: level_offsets must be non-decreasing");
            // count_well_duplicates.py:249 asserts that every ring holds a well -- when it gets to the target, i.e.
            // in a tile where the centre passes the filter: the kernels report that (wd_count_fetch)
            if (e == s) has_empty = true;
            level_len[(size_t)i * levels + l] = e - s;
        }
        for (uint32_t k = a; k < b; ++k) tmp.emplace_back(idx[k], k);
        std::stable_sort(tmp.begin(), tmp.end(),
                         [](const std::pair<uint32_t, uint32_t> &x, const std::pair<uint32_t, uint32_t> &y) {
                             return x.first < y.first;
                         });
        for (const auto &w : tmp) {
            // ring of CSR position w.second
            int l = 0;
            while (level_offsets[(size_t)i * levels + l + 1] <= w.second) ++l;
            slot_well[pos] = w.first;
            slot_level[pos] = (uint8_t)(l + 1);
            slot_csr[pos] = w.second;
            max_well = std::max(max_well, w.first);
            ++pos;
        }
    }
    tgt_off[t] = pos;
    std::vector<uint32_t> visit(t);
    for (uint32_t i = 0; i < t; ++i) visit[i] = i;
    std::stable_sort(visit.begin(), visit.end(), [&](uint32_t x, uint32_t y) { return centres[x] < centres[y]; });
    TargetList &tl = ctx->targets;
    cudaStream_t st = ctx->stream;
    WD_CUDA(cudaStreamSynchronize(st));
    WD_TRY(tl.tgt_off.reserve((size_t)(t + 1) * 4));
    WD_TRY(tl.slot_well.reserve((size_t)n_slots * 4));
    WD_TRY(tl.slot_csr.reserve((size_t)n_slots * 4));
    WD_TRY(tl.slot_level.reserve((size_t)n_slots));
    WD_TRY(tl.level_len.reserve(nseg * 4));
    WD_TRY(tl.visit.reserve((size_t)t * 4));
    WD_CUDA(cudaMemcpyAsync(tl.visit.p, visit.data(), (size_t)t * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(tl.tgt_off.p, tgt_off.data(), (size_t)(t + 1) * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(tl.slot_well.p, slot_well.data(), (size_t)n_slots * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(tl.slot_csr.p, slot_csr.data(), (size_t)n_slots * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(tl.slot_level.p, slot_level.data(), (size_t)n_slots, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemcpyAsync(tl.level_len.p, level_len.data(), nseg * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaStreamSynchronize(st));
    tl.t = t;
    tl.levels = levels;
    tl.n_slots = n_slots;
    tl.max_well = max_well;
    tl.h_idx.assign(idx, idx + n_ring);
    tl.h_slot_csr.swap(slot_csr);
    tl.has_empty_ring = has_empty;
    return WD_OK;
}

// ---- tile staging ------------------------------------------------------------------------
int wd_tile_begin(wd_ctx *ctx, int tile_slot, uint32_t n_clusters, int n_planes) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_tile_begin: null context");
    if (tile_slot < 0 || tile_slot > 65535) WD_FAIL(WD_E_ARG, "wd_tile_begin: slot %d outside 0..65535", tile_slot);
    if (n_clusters == 0) WD_FAIL(WD_E_ARG, "wd_tile_begin: tile has no clusters");
    if (n_planes < 0 || n_planes > 65536) WD_FAIL(WD_E_ARG, "wd_tile_begin: bad plane count %d", n_planes);
    WD_CUDA(cudaSetDevice(ctx->device));
    if ((size_t)tile_slot >= ctx->slots.size()) ctx->slots.resize((size_t)tile_slot + 1);
    TileSlot &s = ctx->slots[tile_slot];
    const size_t stride = ((size_t)n_clusters + 255) & ~(size_t)255;
    const size_t need = stride * (size_t)std::max(n_planes, 1);
    if (need > s.planes.cap || ((size_t)n_clusters + 256) > s.filter.cap) WD_CUDA(cudaStreamSynchronize(ctx->stream));
    WD_TRY(s.planes.reserve(need));
    WD_TRY(s.filter.reserve(stride));
    s.n = n_clusters;
    s.n_planes = n_planes;
    s.stride = stride;
    s.kind.assign((size_t)n_planes, WD_PLANE_EMPTY);
    s.n_block.assign((size_t)n_planes, 0);
    s.filter_set = false;
    s.rank_valid = false;
    s.kind_dirty = true;
    s.has_excl = false;
    s.mapped = nullptr;
    s.mapped_host = nullptr;
    s.mapped_filter = nullptr;
    s.mapped_filter_host = nullptr;
    return WD_OK;
}

int wd_tile_map_host(wd_ctx *ctx, int tile_slot, uint32_t n_clusters, int n_planes, const uint8_t *planes,
                     size_t stride_bytes, const uint8_t *kinds, const uint32_t *n_block, const uint8_t *filter) {
    if (ctx == nullptr || planes == nullptr) WD_FAIL(WD_E_ARG, "wd_tile_map_host: null argument");
    if (tile_slot < 0 || tile_slot > 65535) WD_FAIL(WD_E_ARG, "wd_tile_map_host: slot %d outside 0..65535", tile_slot);
    if (n_clusters == 0) WD_FAIL(WD_E_ARG, "wd_tile_map_host: tile has no clusters");
    if (n_planes < 1 || n_planes > 65536) WD_FAIL(WD_E_ARG, "wd_tile_map_host: bad plane count %d", n_planes);
    WD_CUDA(cudaSetDevice(ctx->device));
    // the kernels dereference the host block directly: it has to be page-locked and mapped
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, planes);
    if (e != cudaSuccess || attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) {
        cudaGetLastError();
        WD_FAIL(WD_E_ARG, "wd_tile_map_host: planes must live in pinned host memory from wd_host_alloc()");
    }
    cudaPointerAttributes fattr;
    if (filter != nullptr) {
        e = cudaPointerGetAttributes(&fattr, filter);
        if (e != cudaSuccess || fattr.type != cudaMemoryTypeHost || fattr.devicePointer == nullptr) {
            cudaGetLastError();
            WD_FAIL(WD_E_ARG, "wd_tile_map_host: filter must live in pinned host memory from wd_host_alloc()");
        }
    }
    for (int p = 0; p < n_planes; ++p) {
        const int k = kinds ? kinds[p] : WD_PLANE_BCL;
        if (k != WD_PLANE_BCL && k != WD_PLANE_CBCL && k != WD_PLANE_CBCL_EXCL)
            WD_FAIL(WD_E_ARG, "wd_tile_map_host: plane %d has kind %d", p, k);
        const uint64_t nb = (k == WD_PLANE_BCL || n_block == nullptr) ? n_clusters : n_block[p];
        const uint64_t bytes = k == WD_PLANE_BCL ? nb : (nb + 1) / 2;
        if (bytes > stride_bytes) WD_FAIL(WD_E_ASSERT, "wd_tile_map_host: plane %d needs %llu bytes, stride is %zu", p,
                                          (unsigned long long)bytes, stride_bytes);
        // same checks as wd_tile_put_cbcl (cbcl_read.py:130-133)
        if (k == WD_PLANE_CBCL && nb != n_clusters)
            WD_FAIL(WD_E_ASSERT, "CBCL block holds %llu clusters, filter says %u (cbcl_read.py:133)", (unsigned long long)nb, n_clusters);
        if (k == WD_PLANE_CBCL_EXCL && nb > n_clusters)
            WD_FAIL(WD_E_ASSERT, "excluded CBCL block holds %llu clusters, tile has only %u", (unsigned long long)nb, n_clusters);
    }
    if ((size_t)tile_slot >= ctx->slots.size()) ctx->slots.resize((size_t)tile_slot + 1);
    TileSlot &s = ctx->slots[tile_slot];
    const size_t fstride = ((size_t)n_clusters + 255) & ~(size_t)255;
    if (fstride > s.filter.cap) WD_CUDA(cudaStreamSynchronize(ctx->stream));
    WD_TRY(s.filter.reserve(fstride));
    s.n = n_clusters;
    s.n_planes = n_planes;
    s.stride = stride_bytes;
    s.kind.resize((size_t)n_planes);
    s.n_block.resize((size_t)n_planes);
    s.has_excl = false;
    for (int p = 0; p < n_planes; ++p) {
        s.kind[p] = kinds ? kinds[p] : (uint8_t)WD_PLANE_BCL;
        s.n_block[p] = (s.kind[p] == WD_PLANE_BCL || n_block == nullptr) ? n_clusters : n_block[p];
        if (s.kind[p] == WD_PLANE_CBCL_EXCL) s.has_excl = true;
    }
    s.filter_set = filter != nullptr;
    s.rank_valid = false;
    s.kind_dirty = true;
    s.mapped = static_cast<const uint8_t *>(attr.devicePointer);
    s.mapped_host = planes;
    s.mapped_filter = filter ? static_cast<const uint8_t *>(fattr.devicePointer) : nullptr;
    s.mapped_filter_host = filter;
    return WD_OK;
}

int wd_tile_put_filter(wd_ctx *ctx, int tile_slot, const uint8_t *bytes, uint32_t n) {
    WD_TRY(check_slot(ctx, tile_slot, "wd_tile_put_filter"));
    TileSlot &s = ctx->slots[tile_slot];
    // bcl_direct_reader.py:236: the filter must describe exactly this tile
    if (n != s.n) WD_FAIL(WD_E_ASSERT, "filter holds %u clusters, tile has %u", n, s.n);
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaMemcpyAsync(s.filter.p, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    // K3 reads whole 64-byte blocks: zero the tail (the plane stride of a mapped slot says nothing about it)
    const size_t fstride = ((size_t)n + 255) & ~(size_t)255;
    if (fstride > n) WD_CUDA(cudaMemsetAsync(s.filter.as<uint8_t>() + n, 0, fstride - n, ctx->stream));
    s.filter_set = true;
    s.rank_valid = false;
    s.mapped_filter = nullptr;
    s.mapped_filter_host = nullptr;
    return WD_OK;
}

int wd_tile_put_bcl(wd_ctx *ctx, int tile_slot, int plane, const uint8_t *bytes, uint32_t n) {
    WD_TRY(check_slot(ctx, tile_slot, "wd_tile_put_bcl"));
    TileSlot &s = ctx->slots[tile_slot];
    if (plane < 0 || plane >= s.n_planes) WD_FAIL(WD_E_ARG, "wd_tile_put_bcl: plane %d outside 0..%d", plane, s.n_planes - 1);
    if (s.mapped) WD_FAIL(WD_E_ARG, "wd_tile_put_bcl: slot %d is mapped to host memory (wd_tile_map_host)", tile_slot);
    // bcl_direct_reader.py:333-338
    if (n != s.n) WD_FAIL(WD_E_ASSERT, "BCL header says %u clusters, filter says %u", n, s.n);
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaMemcpyAsync(s.planes.as<uint8_t>() + (size_t)plane * s.stride, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    if (s.kind[plane] != WD_PLANE_BCL) s.kind_dirty = true;
    s.kind[plane] = WD_PLANE_BCL;
    s.n_block[plane] = n;
    return WD_OK;
}

int wd_tile_put_cbcl(wd_ctx *ctx, int tile_slot, int plane, const uint8_t *nibbles, uint32_t usize, uint32_t n_block,
                     int excluded) {
    WD_TRY(check_slot(ctx, tile_slot, "wd_tile_put_cbcl"));
    TileSlot &s = ctx->slots[tile_slot];
    if (plane < 0 || plane >= s.n_planes) WD_FAIL(WD_E_ARG, "wd_tile_put_cbcl: plane %d outside 0..%d", plane, s.n_planes - 1);
    if (s.mapped) WD_FAIL(WD_E_ARG, "wd_tile_put_cbcl: slot %d is mapped to host memory (wd_tile_map_host)", tile_slot);
    if ((uint64_t)usize * 2 < n_block) WD_FAIL(WD_E_ASSERT, "CBCL block of %u bytes cannot hold %u clusters", usize, n_block);
    if (!excluded && n_block != s.n)
        WD_FAIL(WD_E_ASSERT, "CBCL block holds %u clusters, filter says %u (cbcl_read.py:133)", n_block, s.n);
    if (excluded && n_block > s.n) WD_FAIL(WD_E_ASSERT, "excluded CBCL block holds %u clusters, tile has only %u", n_block, s.n);
    const uint32_t used = (n_block + 1) / 2;
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_CUDA(cudaMemcpyAsync(s.planes.as<uint8_t>() + (size_t)plane * s.stride, nibbles, used, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t k = excluded ? WD_PLANE_CBCL_EXCL : WD_PLANE_CBCL;
    if (s.kind[plane] != k) s.kind_dirty = true;
    s.kind[plane] = k;
    s.n_block[plane] = n_block;
    if (excluded) s.has_excl = true;
    return WD_OK;
}

int wd_filter_offsets(wd_ctx *ctx, int tile_slot, int32_t *offsets, uint32_t *passing) {
    WD_TRY(check_slot(ctx, tile_slot, "wd_filter_offsets"));
    if (offsets == nullptr) WD_FAIL(WD_E_ARG, "wd_filter_offsets: null output");
    WD_CUDA(cudaSetDevice(ctx->device));
    return filter_offsets(ctx, tile_slot, offsets, passing);
}

int wd_get_seqs(wd_ctx *ctx, int tile_slot, const int64_t *indices, uint32_t n_idx, const int32_t *plane_order,
                int seq_len, uint8_t *codes, uint8_t *pf) {
    if (ctx == nullptr || (n_idx && (indices == nullptr || pf == nullptr)) || (seq_len > 0 && n_idx && (plane_order == nullptr || codes == nullptr)))
        WD_FAIL(WD_E_ARG, "wd_get_seqs: null argument");
    if (seq_len < 0 || seq_len > WD_MAX_SEQ_LEN) WD_FAIL(WD_E_ARG, "wd_get_seqs: sequence length %d outside 0..%d", seq_len, WD_MAX_SEQ_LEN);
    WD_CUDA(cudaSetDevice(ctx->device));
    return get_seqs(ctx, tile_slot, indices, n_idx, plane_order, seq_len, codes, pf);
}

// ---- stage 3 ---------------------------------------------------------------------------------
int wd_count_async(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order, int seq_len,
                   int edit_distance, int hamming, int mode, int want_per_target) {
    if (ctx == nullptr || plane_order == nullptr) WD_FAIL(WD_E_ARG, "wd_count: null argument");
    WD_CUDA(cudaSetDevice(ctx->device));
    return count_async(ctx, first_slot, n_tiles, plane_order, seq_len, edit_distance, hamming, mode, want_per_target);
}

int wd_count_fetch(wd_ctx *ctx, int32_t *per_target, int64_t *tile_counters) {
    if (ctx == nullptr) WD_FAIL(WD_E_ARG, "wd_count_fetch: null context");
    if (ctx->last_tiles == 0) WD_FAIL(WD_E_ARG, "wd_count_fetch: no count has been issued");
    WD_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t width = 1 + 5 * (size_t)ctx->last_levels;
    const size_t n_cnt = (size_t)ctx->last_tiles * width;
    if (tile_counters) WD_CUDA(cudaMemcpyAsync(tile_counters, ctx->counters.p, n_cnt * 8, cudaMemcpyDeviceToHost, st));
    if (per_target) {
        if (!ctx->last_per_target) WD_FAIL(WD_E_ARG, "wd_count_fetch: the count was issued without per-target output");
        WD_CUDA(cudaMemcpyAsync(per_target, ctx->per_target.p,
                                (size_t)ctx->last_tiles * ctx->last_t * (1 + 2 * (size_t)ctx->last_levels) * 4,
                                cudaMemcpyDeviceToHost, st));
    }
    // two status words behind the counter rows, written by the counting kernels
    unsigned long long status[2] = {0, 0};
    WD_CUDA(cudaMemcpyAsync(status, ctx->counters.as<unsigned long long>() + n_cnt, 16, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    if (status[1] != 0) {
        // excluded CBCL blocks must hold exactly the PF wells (cbcl_read.py:130-131)
        const int k = (int)(uint32_t)~status[1];
        TileSlot &s = ctx->slots[ctx->last_first_slot + k];
        uint32_t pf = 0;
        WD_CUDA(cudaMemcpy(&pf, s.pfrank.as<uint32_t>() + (s.n + 63) / 64, 4, cudaMemcpyDeviceToHost));
        for (int p = 0; p < s.n_planes; ++p)
            if (s.kind[p] == WD_PLANE_CBCL_EXCL && s.n_block[p] != pf)
                WD_FAIL(WD_E_ASSERT, "tile slot %d plane %d: excluded CBCL block holds %u clusters but %u wells pass the filter",
                        ctx->last_first_slot + k, p, s.n_block[p], pf);
    }
    if (status[0] != 0) {
        const unsigned long long v = ~status[0];
        WD_FAIL(WD_E_ASSERT, "target %u has a ring without wells (its centre passes the filter of tile %u of the batch)",
                (uint32_t)v, (uint32_t)(v >> 32));
    }
    return WD_OK;
}

int wd_count(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order, int seq_len, int edit_distance,
             int hamming, int mode, int32_t *per_target, int64_t *tile_counters) {
    WD_TRY(wd_count_async(ctx, first_slot, n_tiles, plane_order, seq_len, edit_distance, hamming, mode,
                          per_target != nullptr));
    return wd_count_fetch(ctx, per_target, tile_counters);
}

int wd_count_trace_sectors(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order, int seq_len,
                           int edit_distance, int hamming, uint32_t *sectors, uint32_t *lines) {
    if (ctx == nullptr || plane_order == nullptr || sectors == nullptr || lines == nullptr)
        WD_FAIL(WD_E_ARG, "wd_count_trace_sectors: null argument");
    WD_CUDA(cudaSetDevice(ctx->device));
    return count_trace(ctx, first_slot, n_tiles, plane_order, seq_len, edit_distance, hamming, sectors, lines);
}

// reference log order: tile, then target / level / well in file order == position in the caller's list
static int dup_pairs_sorted(wd_ctx *ctx, std::vector<int32_t> &raw, std::vector<size_t> &ord) {
    if (ctx->dup_cap == 0 || ctx->last_tiles == 0)
        WD_FAIL(WD_E_ARG, "wd_dup_pairs: the last count did not log duplicate pairs (mode 1 or 2)");
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_TRY(dup_rows_fetch(ctx, raw));
    const size_t n = raw.size() / 4;
    const std::vector<uint32_t> &csr = ctx->targets.h_slot_csr;
    ord.resize(n);
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
        if (raw[a * 4] != raw[b * 4]) return raw[a * 4] < raw[b * 4];
        return csr[(uint32_t)raw[a * 4 + 2]] < csr[(uint32_t)raw[b * 4 + 2]];
    });
    return WD_OK;
}

int wd_dup_pairs_seqs(wd_ctx *ctx, int32_t *rows, uint8_t *codes, size_t cap, uint64_t *n_rows) {
    if (ctx == nullptr || n_rows == nullptr) WD_FAIL(WD_E_ARG, "wd_dup_pairs: null argument");
    std::vector<int32_t> raw;
    std::vector<size_t> ord;
    WD_TRY(dup_pairs_sorted(ctx, raw, ord));
    const size_t n = ord.size();
    *n_rows = n;
    if (n == 0) return WD_OK;
    if (n > cap || rows == nullptr) WD_FAIL(WD_E_CAPACITY, "wd_dup_pairs: %zu rows needed, caller provided %zu", n, cap);
    std::vector<uint8_t> h_codes;
    if (codes != nullptr) WD_TRY(dup_seqs(ctx, raw, h_codes));
    const size_t len2 = 2 * (size_t)ctx->last_seq_len;
    for (size_t i = 0; i < n; ++i) {
        const int32_t *r = &raw[ord[i] * 4];
        rows[i * 4 + 0] = r[0];
        rows[i * 4 + 1] = r[1];
        rows[i * 4 + 2] = (int32_t)ctx->targets.h_idx[ctx->targets.h_slot_csr[(uint32_t)r[2]]];
        rows[i * 4 + 3] = r[3];
        if (codes != nullptr) memcpy(codes + i * len2, h_codes.data() + ord[i] * len2, len2);
    }
    return WD_OK;
}

int wd_dup_pairs(wd_ctx *ctx, int32_t *rows, size_t cap, uint64_t *n_rows) {
    return wd_dup_pairs_seqs(ctx, rows, nullptr, cap, n_rows);
}

// ---- multi-GPU -----------------------------------------------------------------------------------
static int publish_any(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles, int n_rows_total,
                       void **devptr, size_t *n_int64, bool keep) {
    if (ctx == nullptr || tile_row == nullptr || lane_row == nullptr) WD_FAIL(WD_E_ARG, "wd_publish_counters: null argument");
    if (n_rows_total < 1) WD_FAIL(WD_E_ARG, "wd_publish_counters: n_rows_total must be positive");
    WD_CUDA(cudaSetDevice(ctx->device));
    WD_TRY(publish_counters(ctx, tile_row, lane_row, n_tiles, n_rows_total, keep));
    if (devptr) *devptr = ctx->publish.as<unsigned long long>() + (size_t)ctx->publish_cur * ctx->publish_n;
    if (n_int64) *n_int64 = ctx->publish_n;
    return WD_OK;
}

int wd_publish_counters(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles, int n_rows_total,
                        void **devptr, size_t *n_int64) {
    return publish_any(ctx, tile_row, lane_row, n_tiles, n_rows_total, devptr, n_int64, false);
}

int wd_publish_add(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles, int n_rows_total,
                   void **devptr, size_t *n_int64) {
    return publish_any(ctx, tile_row, lane_row, n_tiles, n_rows_total, devptr, n_int64, true);
}

int wd_counters_devptr(wd_ctx *ctx, void **devptr, size_t *n_int64) {
    if (ctx == nullptr || devptr == nullptr || n_int64 == nullptr) WD_FAIL(WD_E_ARG, "wd_counters_devptr: null argument");
    if (ctx->last_tiles == 0) WD_FAIL(WD_E_ARG, "wd_counters_devptr: no count has been issued");
    *devptr = ctx->counters.p;
    *n_int64 = (size_t)ctx->last_tiles * (1 + 5 * (size_t)ctx->last_levels);
    return WD_OK;
}

int wd_count_exhaustive(wd_ctx *ctx, int tile_slot, const int32_t *plane_order, int seq_len, int levels,
                        uint32_t window_lo, uint32_t window_hi, int edit_distance, int hamming, int64_t *tile_counters) {
    WD_TRY(check_slot(ctx, tile_slot, "wd_count_exhaustive"));
    if (plane_order == nullptr || tile_counters == nullptr) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: null argument");
    WD_CUDA(cudaSetDevice(ctx->device));
    return count_exhaustive(ctx, tile_slot, plane_order, seq_len, levels, window_lo, window_hi, edit_distance, hamming,
                            tile_counters);
}

}  // extern "C"
