// Stage 1 -- hex-lattice neighbourhood construction (K0, K1, K2).
//
// Replaces the per-target linear scan of prepare_cluster_indexes.py:38-78 by a
// uniform-grid spatial hash: wells are binned into cells 32 px wide and 128 px
// tall (wd_common.cuh), and each target is served by one warp that tests only
// the wells of three runs of cells, one per grid row around the centre.
// The membership rule is the reference's, restated in integers:
//   MAX[l] < sqrt(dx^2+dy^2) <= MAX[l+1]   <=>   MAX[l]^2 < dx^2+dy^2 <= MAX[l+1]^2
// (exact: sqrt is correctly rounded and the thresholds are integers), together
// with the reference's index window [c-20000, c+20001] (:52-67).
#include "wd_common.cuh"
#include "wd_scan.cuh"

#include <limits.h>

namespace wd {

constexpr int RING_LEVELS = 5;                      // len(MAX_DISTS) - 1, prepare_cluster_indexes.py:19
__constant__ int c_ring_d2[RING_LEVELS + 1] = {1, 484, 1764, 3844, 6724, 10404};   // MAX_DISTS^2

// ---- K0: .locs floats -> integer pixels -------------------------------------------
// x = int(f * 10.0 + 1000.5) evaluated in float64 with truncation toward zero
// (prepare_cluster_indexes.py:110-112).  bbox = {min x, min y, max x, max y}.
__global__ void __launch_bounds__(256)
locs_to_pixels_kernel(const float2 *__restrict__ xy, uint32_t n, int *__restrict__ px,
                      int *__restrict__ py, int *__restrict__ bbox, int *__restrict__ bad) {
    int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 f = xy[i];
        const double vx = __dadd_rn(__dmul_rn((double)f.x, 10.0), 1000.5);
        const double vy = __dadd_rn(__dmul_rn((double)f.y, 10.0), 1000.5);
        // Python's int() raises on nan/inf and has no range limit; pixels beyond +-2^30 are rejected
        if (!(fabs(vx) < 1073741824.0) || !(fabs(vy) < 1073741824.0)) {
            atomicExch(bad, 1);
            px[i] = 0;
            py[i] = 0;
            continue;
        }
        const int ix = (int)__double2ll_rz(vx);
        const int iy = (int)__double2ll_rz(vy);
        px[i] = ix;
        py[i] = iy;
        mnx = min(mnx, ix); mny = min(mny, iy); mxx = max(mxx, ix); mxy = max(mxy, iy);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], mnx); atomicMin(&bbox[1], mny);
        atomicMax(&bbox[2], mxx); atomicMax(&bbox[3], mxy);
    }
}

// ---- K1: uniform grid (histogram -> scan -> scatter) -------------------------------------
__device__ __forceinline__ uint32_t cell_of(int x, int y, int min_x, int min_y, int grid_w) {
    return (uint32_t)((y - min_y) >> CELL_SHIFT_Y) * (uint32_t)grid_w + (uint32_t)((x - min_x) >> CELL_SHIFT_X);
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const int *__restrict__ px, const int *__restrict__ py, uint32_t n, int min_x,
                  int min_y, int grid_w, uint32_t *__restrict__ counts) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&counts[cell_of(px[i], py[i], min_x, min_y, grid_w)], 1u);
}

// cell_wells[pos] = {x, y, well index, 0}: one 16-byte record per well, cell after cell
__global__ void __launch_bounds__(256)
cell_scatter_kernel(const int *__restrict__ px, const int *__restrict__ py, uint32_t n, int min_x,
                    int min_y, int grid_w, const uint32_t *__restrict__ cell_start,
                    uint32_t *__restrict__ cursor, int4 *__restrict__ cell_wells) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = px[i], y = py[i];
        const uint32_t c = cell_of(x, y, min_x, min_y, grid_w);
        const uint32_t pos = cell_start[c] + atomicAdd(&cursor[c], 1u);
        cell_wells[pos] = make_int4(x, y, (int)i, 0);
    }
}

// ---- K2: ring query, one warp per target -------------------------------------------------
struct RingArgs {
    const int *px, *py;
    const uint32_t *cell_start;
    const int4 *cell_wells;
    const uint32_t *centres;
    uint32_t n, t;
    int levels, min_x, min_y, grid_w, grid_h;
    uint32_t wlo, whi;
};

// FILL = false: counts[t*levels + l] and first_empty.  FILL = true: write the
// wells of each ring into out[] at offsets[], ascending.
template <bool FILL>
__global__ void __launch_bounds__(256)
ring_query_kernel(RingArgs a, uint32_t *__restrict__ counts, uint32_t *__restrict__ first_empty,
                  const uint32_t *__restrict__ offsets, uint32_t *__restrict__ tmp,
                  uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.t) return;
    const uint32_t c = a.centres[t];
    const int cx = a.px[c], cy = a.py[c];
    const long long lo = (long long)c - (long long)a.wlo;
    const long long hi = (long long)c + (long long)a.whi;
    const int gy = (cy - a.min_y) >> CELL_SHIFT_Y;
    const int x0 = max(cx - RING_RADIUS - a.min_x, 0) >> CELL_SHIFT_X;
    const int x1 = min((cx + RING_RADIUS - a.min_x) >> CELL_SHIFT_X, a.grid_w - 1);
    uint32_t run[RING_LEVELS];
#pragma unroll
    for (int l = 0; l < RING_LEVELS; ++l) run[l] = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int yy = max(gy - 1, 0); yy <= min(gy + 1, a.grid_h - 1); ++yy) {
        // the cells of one grid row are contiguous in cell_wells
        const uint32_t s = a.cell_start[(uint32_t)yy * a.grid_w + x0];
        const uint32_t e = a.cell_start[(uint32_t)yy * a.grid_w + x1 + 1];
        for (uint32_t base = s; base < e; base += 32) {
            const uint32_t i = base + lane;
            int lvl = -1;
            int widx = 0;
            if (i < e) {
                const int4 w = a.cell_wells[i];
                const int dx = w.x - cx, dy = w.y - cy;
                widx = w.z;
                if (abs(dx) <= RING_RADIUS && abs(dy) <= RING_RADIUS && (long long)w.z >= lo && (long long)w.z <= hi) {
                    const int d2 = dx * dx + dy * dy;
#pragma unroll
                    for (int l = 0; l < RING_LEVELS; ++l)
                        if (d2 > c_ring_d2[l] && d2 <= c_ring_d2[l + 1]) lvl = l;
                }
            }
#pragma unroll
            for (int l = 0; l < RING_LEVELS; ++l) {
                if (l < a.levels) {
                    const uint32_t m = __ballot_sync(0xffffffffu, lvl == l);
                    if (FILL && lvl == l)
                        tmp[offsets[(size_t)t * a.levels + l] + run[l] + __popc(m & lt_mask)] = (uint32_t)widx;
                    run[l] += __popc(m);
                }
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int l = 0; l < RING_LEVELS; ++l) {
            if (l < a.levels && lane == l) {
                counts[(size_t)t * a.levels + l] = run[l];
                if (run[l] == 0) atomicMin(first_empty, t * (uint32_t)a.levels + l);
            }
        }
    } else {
        // scan order inside a ring: ascending well index (prepare_cluster_indexes.py:56-63)
        __syncwarp();
#pragma unroll
        for (int l = 0; l < RING_LEVELS; ++l) {
            if (l < a.levels) {
                const uint32_t off = offsets[(size_t)t * a.levels + l];
                const uint32_t m = run[l];
                for (uint32_t i = lane; i < m; i += 32) {
                    const uint32_t v = tmp[off + i];
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < m; ++j) rank += tmp[off + j] < v;
                    out[off + rank] = v;
                }
            }
        }
    }
}

static int grid_for(size_t n, int threads, int sm_count) {
    size_t b = (n + threads - 1) / threads;
    size_t cap = (size_t)sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int locs_load(wd_ctx *ctx, const float *xy, uint32_t n) {
    if (n == 0) WD_FAIL(WD_E_ARG, "wd_locs_load: empty .locs");
    cudaStream_t st = ctx->stream;
    WD_TRY(ctx->xy.reserve((size_t)n * 8));
    WD_TRY(ctx->px.reserve((size_t)n * 4));
    WD_TRY(ctx->py.reserve((size_t)n * 4));
    WD_TRY(ctx->bbox.reserve(8 * sizeof(int)));
    WD_CUDA(cudaMemcpyAsync(ctx->xy.p, xy, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    const int init[8] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN, 0, 0, 0, 0};
    // pageable source: staged by the runtime before the call returns
    WD_CUDA(cudaMemcpyAsync(ctx->bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    int *bbox = ctx->bbox.as<int>();
    locs_to_pixels_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(
        ctx->xy.as<float2>(), n, ctx->px.as<int>(), ctx->py.as<int>(), bbox, bbox + 4);
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    int h[8];
    WD_CUDA(cudaMemcpyAsync(h, bbox, sizeof(h), cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    if (h[4]) WD_FAIL(WD_E_ARG, "wd_locs_load: non-finite or out-of-range coordinate in .locs");
    ctx->min_x = h[0];
    ctx->min_y = h[1];
    const long long gw = (((long long)h[2] - h[0]) >> CELL_SHIFT_X) + 1;
    const long long gh = (((long long)h[3] - h[1]) >> CELL_SHIFT_Y) + 1;
    if (gw * gh > (1ll << 26))
        WD_FAIL(WD_E_ARG, "wd_locs_load: coordinate bounding box %lld x %lld cells is too large", gw, gh);
    ctx->grid_w = (int)gw;
    ctx->grid_h = (int)gh;
    const size_t n_cells = (size_t)(gw * gh);
    WD_TRY(ctx->cell_start.reserve((n_cells + 1) * 4));
    WD_TRY(ctx->cell_cursor.reserve(n_cells * 4));
    WD_TRY(ctx->cell_wells.reserve((size_t)n * 16));
    WD_TRY(ctx->scan_tmp.reserve(scan_tmp_words(n_cells) * 4));
    WD_CUDA(cudaMemsetAsync(ctx->cell_cursor.p, 0, n_cells * 4, st));
    const int g = grid_for(n, 256, ctx->sm_count);
    cell_count_kernel<<<g, 256, 0, st>>>(ctx->px.as<int>(), ctx->py.as<int>(), n, ctx->min_x, ctx->min_y,
                                         ctx->grid_w, ctx->cell_cursor.as<uint32_t>());
    ctx->launches++;
    WD_CUDA(exclusive_scan_u32(ctx->cell_cursor.as<uint32_t>(), ctx->cell_start.as<uint32_t>(), n_cells,
                               ctx->scan_tmp.as<uint32_t>(), st, &ctx->launches));
    WD_CUDA(cudaMemsetAsync(ctx->cell_cursor.p, 0, n_cells * 4, st));
    cell_scatter_kernel<<<g, 256, 0, st>>>(ctx->px.as<int>(), ctx->py.as<int>(), n, ctx->min_x, ctx->min_y,
                                           ctx->grid_w, ctx->cell_start.as<uint32_t>(),
                                           ctx->cell_cursor.as<uint32_t>(), ctx->cell_wells.as<int4>());
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    ctx->n_locs = n;
    ctx->q_t = 0;
    ctx->x_geom_levels = 0;          // ring sizes of exhaustive mode belong to the previous .locs
    return WD_OK;
}

int ring_query(wd_ctx *ctx, const uint32_t *centres, uint32_t t, int levels, uint32_t wlo,
               uint32_t whi, uint32_t *level_offsets, uint32_t *idx, size_t idx_cap,
               uint64_t *n_idx, uint32_t *first_empty) {
    if (ctx->n_locs == 0) WD_FAIL(WD_E_ARG, "wd_ring_query: call wd_locs_load first");
    if (levels < 1 || levels > RING_LEVELS)
        WD_FAIL(WD_E_ARG, "wd_ring_query: levels must be 1..%d (MAX_DISTS defines %d rings)", RING_LEVELS, RING_LEVELS);
    if (first_empty) *first_empty = UINT32_MAX;
    if (n_idx) *n_idx = 0;
    for (uint32_t i = 0; i < t; ++i)
        if (centres[i] >= ctx->n_locs)
            WD_FAIL(WD_E_INDEX, "wd_ring_query: centre %u is out of range (tile has %u wells)", centres[i], ctx->n_locs);
    if (t == 0) {
        level_offsets[0] = 0;
        return WD_OK;
    }
    cudaStream_t st = ctx->stream;
    const size_t nseg = (size_t)t * levels;
    WD_TRY(ctx->q_centres.reserve((size_t)t * 4));
    WD_TRY(ctx->q_counts.reserve(nseg * 4));
    WD_TRY(ctx->q_offsets.reserve((nseg + 1) * 4));
    WD_TRY(ctx->q_flag.reserve(4));
    WD_TRY(ctx->scan_tmp.reserve(scan_tmp_words(nseg) * 4));
    WD_CUDA(cudaMemcpyAsync(ctx->q_centres.p, centres, (size_t)t * 4, cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaMemsetAsync(ctx->q_flag.p, 0xff, 4, st));
    RingArgs a;
    a.px = ctx->px.as<int>(); a.py = ctx->py.as<int>();
    a.cell_start = ctx->cell_start.as<uint32_t>(); a.cell_wells = ctx->cell_wells.as<int4>();
    a.centres = ctx->q_centres.as<uint32_t>();
    a.n = ctx->n_locs; a.t = t; a.levels = levels;
    a.min_x = ctx->min_x; a.min_y = ctx->min_y; a.grid_w = ctx->grid_w; a.grid_h = ctx->grid_h;
    a.wlo = wlo; a.whi = whi;
    const unsigned blocks = (t + 7) / 8;
    ring_query_kernel<false><<<blocks, 256, 0, st>>>(a, ctx->q_counts.as<uint32_t>(), ctx->q_flag.as<uint32_t>(),
                                                     nullptr, nullptr, nullptr);
    ctx->launches++;
    WD_CUDA(exclusive_scan_u32(ctx->q_counts.as<uint32_t>(), ctx->q_offsets.as<uint32_t>(), nseg,
                               ctx->scan_tmp.as<uint32_t>(), st, &ctx->launches));
    uint32_t fe = UINT32_MAX;
    WD_CUDA(cudaMemcpyAsync(level_offsets, ctx->q_offsets.p, (nseg + 1) * 4, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaMemcpyAsync(&fe, ctx->q_flag.p, 4, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    const uint64_t total = level_offsets[nseg];
    if (n_idx) *n_idx = total;
    if (first_empty) *first_empty = fe;
    if (fe != UINT32_MAX) {
        ctx->q_t = 0;
        WD_FAIL(WD_E_RUNTIME, "Got no wells for cluster %u level %u", centres[fe / levels], fe % levels);
    }
    if (total > idx_cap) {
        ctx->q_t = 0;
        WD_FAIL(WD_E_CAPACITY, "wd_ring_query: idx holds %zu entries, %llu needed", idx_cap,
                (unsigned long long)total);
    }
    WD_TRY(ctx->q_idx.reserve((size_t)total * 4 + 4));
    WD_TRY(ctx->q_tmp.reserve((size_t)total * 4 + 4));
    ring_query_kernel<true><<<blocks, 256, 0, st>>>(a, nullptr, nullptr, ctx->q_offsets.as<uint32_t>(),
                                                    ctx->q_tmp.as<uint32_t>(), ctx->q_idx.as<uint32_t>());
    ctx->launches++;
    WD_CUDA(cudaGetLastError());
    if (idx) WD_CUDA(cudaMemcpyAsync(idx, ctx->q_idx.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    ctx->q_t = t;
    ctx->q_levels = levels;
    ctx->q_total = total;
    return WD_OK;
}

}  // namespace wd
