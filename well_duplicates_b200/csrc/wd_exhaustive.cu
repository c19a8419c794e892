// Exhaustive mode (BASELINE config 3): every well of a tile is a target.
//
// No target list exists -- it would be 388 M entries per tile.  The work is
// split by what it depends on:
//
//  * geometry (prepare_cluster_indexes.py:19,32-35,52-67: ring of every well
//    within 102 px under the float64 distance rule, restated in integers, and
//    the index window [c-20000, c+20001]) depends only on the .locs file, which
//    a whole flowcell shares (Snakefile.count_dups:168).  exh_geometry_kernel
//    computes the five ring sizes of every well ONCE per wd_locs_load (the
//    LENGTH column of the reference, and its "Got no wells" RuntimeError,
//    :70-76) and keeps them in HBM;
//  * sequences change per tile.  A dense pass packs every well (coalesced plane
//    reads, wd_kernels23.cuh), exh_prefix_kernel lays the first 32 symbols of
//    every well out in GRID order (the order of stage 1's cell records) and
//    resets the per-well tallies, exh_compare_kernel looks at every pair of
//    wells that could be ring neighbours, exh_verify_kernel runs the exact test
//    on the pairs that passed and exh_finish_kernel sums the tallies into the
//    counters of the report.
//
// The geometry and the compare kernel give one THREAD to each centre and one
// candidate list to each WARP: the 32 centres of a warp are consecutive grid
// records (wells of the same few cells), so one list -- runs of records that
// cover all 32 neighbourhoods -- serves them all.  The warp stages 32
// candidates at a time in shared memory and every lane tests the same
// candidate against its own centre (broadcast LDS, no divergence).  For the
// compare that test is a 32-symbol necessary condition for dist <= e on two
// bit-planes (Head32Sets, wd_seq.cuh: three LOP3 and a POPC); distance and ring
// are symmetric, so every unordered pair is looked at once and a duplicate is
// credited to both wells.  Only the pairs that pass -- real duplicates, about
// one in 10^4 -- are queued for the exact test: ring, both index windows,
// full-length compare on the packed words, one pair per thread.
#include <algorithm>
#include <climits>
#include <cstring>

#include "wd_common.cuh"
#include "wd_pack.cuh"
#include "wd_seq.cuh"

namespace wd {

constexpr int XW = 8;                   // warps per CTA
constexpr int XFIELD = 12;              // bits per ring in a centre's packed tallies
constexpr uint32_t XFIELD_MAX = (1u << XFIELD) - 1u;
__constant__ int c_x_d2[6] = {1, 484, 1764, 3844, 6724, 10404};   // MAX_DISTS^2, prepare_cluster_indexes.py:19

struct XGrid {
    const uint32_t *cell_start;
    const int4 *cell_wells;             // {x, y, well, 0} in grid order
    int min_x, min_y, grid_w, grid_h;
    uint32_t n;
};

// The candidate list of a warp: up to three runs of grid records (one per grid
// row; the cells of a grid row are contiguous) that hold every well within
// RING_RADIUS of any active lane's centre.  All active lanes sit in grid row
// `g`, so their neighbourhoods reach rows g-1 .. g+1 at most.
struct XRuns {
    uint32_t start[3], len[3], total;
    __device__ __forceinline__ uint32_t record(uint32_t v) const {
        if (v < len[0]) return start[0] + v;
        v -= len[0];
        if (v < len[1]) return start[1] + v;
        return start[2] + (v - len[1]);
    }
};

__device__ __forceinline__ XRuns warp_runs(const XGrid &g, bool active, int cx, int cy, int &row0) {
    const int xmin = __reduce_min_sync(0xffffffffu, active ? cx : INT_MAX);
    const int xmax = __reduce_max_sync(0xffffffffu, active ? cx : INT_MIN);
    const int ymin = __reduce_min_sync(0xffffffffu, active ? cy : INT_MAX);
    const int ymax = __reduce_max_sync(0xffffffffu, active ? cy : INT_MIN);
    const int x0 = max(xmin - RING_RADIUS - g.min_x, 0) >> CELL_SHIFT_X;
    const int x1 = min((xmax + RING_RADIUS - g.min_x) >> CELL_SHIFT_X, g.grid_w - 1);
    const int y0 = max(ymin - RING_RADIUS - g.min_y, 0) >> CELL_SHIFT_Y;
    const int y1 = min((ymax + RING_RADIUS - g.min_y) >> CELL_SHIFT_Y, g.grid_h - 1);
    XRuns r;
    r.total = 0;
    row0 = y0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int yy = y0 + k;
        r.start[k] = 0;
        r.len[k] = 0;
        if (yy <= y1) {
            const uint32_t s = __ldg(g.cell_start + (uint32_t)yy * g.grid_w + x0);
            const uint32_t e = __ldg(g.cell_start + (uint32_t)yy * g.grid_w + x1 + 1);
            r.start[k] = s;
            r.len[k] = e - s;
        }
        r.total += r.len[k];
    }
    return r;
}

// ring (0-based) of a candidate at squared distance d2, or -1:
//   MAX[l] < dist <= MAX[l+1]  <=>  MAX[l]^2 < d2 <= MAX[l+1]^2   (wd_stage1.cu)
__device__ __forceinline__ bool in_rings(int d2, int d2_max) { return (unsigned)(d2 - 2) <= (unsigned)(d2_max - 2); }
__device__ __forceinline__ int ring_of(int d2) { return (d2 > 484) + (d2 > 1764) + (d2 > 3844) + (d2 > 6724); }

// index window [c - wlo, c + whi] as (lo, span): w is inside  <=>  (uint32)(w - lo) <= span
__device__ __forceinline__ void index_window(uint32_t c, uint32_t wlo, uint32_t whi, uint32_t &lo, uint32_t &span) {
    lo = c > wlo ? c - wlo : 0u;
    const unsigned long long hi = min((unsigned long long)c + whi, 0xffffffffull);
    span = (uint32_t)hi - lo;
}

// ============================================================================
// geometry: ring sizes of every well, once per .locs
// ============================================================================
struct XGeomArgs {
    XGrid g;
    unsigned long long *ringlen;         // [n] grid order: ring l in bits [12 l, 12 l + 12)
    uint32_t *first_empty;               // min over (well * levels + ring) of the empty rings
    uint32_t *overflow;
    int levels;
    uint32_t wlo, whi;
};

__global__ void __launch_bounds__(XW * 32)
exh_geometry_kernel(XGeomArgs a) {
    __shared__ int4 s_rec[XW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d2_max = c_x_d2[a.levels];
    for (uint32_t p0 = (blockIdx.x * XW + warp) * 32u; p0 < a.g.n; p0 += gridDim.x * XW * 32u) {
        const uint32_t p = p0 + lane;
        const bool has = p < a.g.n;
        const int4 me = has ? __ldg(a.g.cell_wells + p) : make_int4(0, 0, 0, 0);
        uint32_t lo, span;
        index_window((uint32_t)me.z, a.wlo, a.whi, lo, span);
        const int gy = (me.y - a.g.min_y) >> CELL_SHIFT_Y;
        unsigned long long cnt = 0;
        bool pending = has, too_many = false;
        for (;;) {                                          // one round per grid row the warp's centres sit in
            const int g = __reduce_min_sync(0xffffffffu, pending ? gy : INT_MAX);
            if (g == INT_MAX) break;
            const bool active = pending && gy == g;
            pending = pending && !active;
            int row0;
            const XRuns r = warp_runs(a.g, active, me.x, me.y, row0);
            too_many = too_many || (active && r.total > XFIELD_MAX);
            for (uint32_t v0 = 0; v0 < r.total; v0 += 32) {
                const uint32_t v = v0 + lane;
                __syncwarp();
                if (v < r.total) s_rec[warp][lane] = __ldg(a.g.cell_wells + r.record(v));
                __syncwarp();
                const int nk = (int)min(32u, r.total - v0);
                if (active) {
#pragma unroll 8
                    for (int k = 0; k < nk; ++k) {
                        const int4 w = s_rec[warp][k];
                        const int dx = w.x - me.x, dy = w.y - me.y;         // |dx|, |dy| < 2^9 + grid slack
                        const int d2 = dx * dx + dy * dy;
                        const bool in = in_rings(d2, d2_max) && (uint32_t)((uint32_t)w.z - lo) <= span;
                        cnt += (unsigned long long)(in ? 1u : 0u) << (XFIELD * ring_of(d2));
                    }
                }
            }
        }
        if (has) {
            a.ringlen[p] = cnt;
            if (too_many) atomicExch(a.overflow, 1u);
            // a ring without wells is the reference's RuntimeError, pass-filter centre or not
            for (int l = 0; l < a.levels; ++l)
                if (((cnt >> (XFIELD * l)) & XFIELD_MAX) == 0)
                    atomicMin(a.first_empty, (uint32_t)me.z * (uint32_t)a.levels + (uint32_t)l);
        }
    }
}

// ============================================================================
// per tile: 32-symbol prefixes in grid order, tallies reset
// ============================================================================
constexpr unsigned long long X_INVALID = 1ull << 63;      // tally flag: the well fails the filter

__global__ void __launch_bounds__(256)
exh_prefix_kernel(const int4 *__restrict__ cell_wells, const uint64_t *__restrict__ packed, int words, uint32_t n,
                  uint2 *__restrict__ pre, unsigned long long *__restrict__ tally) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t well = (uint32_t)__ldg(cell_wells + p).z;
    const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(packed + (size_t)well * words * PACK_STRIDE);
    const ulonglong2 q0 = __ldg(q), q1 = __ldg(q + 1);
    pre[p] = make_uint2((uint32_t)q0.x, (uint32_t)q0.y);        // lo, hi; an N reads as A here (necessary test only)
    tally[p] = (q1.y & 1ull) ? 0ull : X_INVALID;                // count_well_duplicates.py:236-237
}

// ============================================================================
// per tile: compare
// ============================================================================
// dist() and the ring of a pair are symmetric, so every unordered pair is
// tested ONCE, by the well that comes first in grid order: a warp's candidate
// list is the rest of its own grid row (from its first record on, as far as
// 102 px to the right of its last centre) and the run of the next grid row.
// A duplicate pair is credited to both wells' tallies (ring l in bits
// [12 l, 12 l + 12)), each under its own index window; wells that fail the
// filter collect tallies too and are skipped by exh_finish_kernel.
struct XCmpArgs {
    XGrid g;
    const uint2 *pre;                    // [n] grid order
    const uint64_t *packed;              // [n][words][4] well order
    uint2 *work;                         // pairs (grid record of the first well, of the second) for exh_verify_kernel
    uint32_t *work_count;
    uint32_t work_cap;
    int words, e, k;
};

__global__ void __launch_bounds__(XW * 32)
exh_compare_kernel(XCmpArgs a) {
    __shared__ uint2 s_pre[XW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int need_all = 32 - a.e;                          // matched positions a pair needs among the first 32
    for (uint32_t p0 = (blockIdx.x * XW + warp) * 32u; p0 < a.g.n; p0 += gridDim.x * XW * 32u) {
        const uint32_t p = p0 + lane;
        const bool has = p < a.g.n;
        int4 me = make_int4(0, 0, 0, 0);
        Head32Sets c;                                       // the 32-symbol necessary test, wd_seq.cuh
        c.clear();
        if (has) {
            me = __ldg(a.g.cell_wells + p);
            const ulonglong2 q0 = __ldg(reinterpret_cast<const ulonglong2 *>(a.packed + (size_t)(uint32_t)me.z * (a.words * PACK_STRIDE)));
            c.set(q0.x, q0.y, a.k);
        }
        const int gy = (me.y - a.g.min_y) >> CELL_SHIFT_Y;
        bool pending = has;
        for (;;) {                                          // one round per grid row the warp's centres sit in
            const int g = __reduce_min_sync(0xffffffffu, pending ? gy : INT_MAX);
            if (g == INT_MAX) break;
            const bool active = pending && gy == g;
            pending = pending && !active;
            const int first = __ffs(__ballot_sync(0xffffffffu, active)) - 1;
            const int xmin = __reduce_min_sync(0xffffffffu, active ? me.x : INT_MAX);
            const int xmax = __reduce_max_sync(0xffffffffu, active ? me.x : INT_MIN);
            const int ymax = __reduce_max_sync(0xffffffffu, active ? me.y : INT_MIN);
            const int x0 = max(xmin - RING_RADIUS - a.g.min_x, 0) >> CELL_SHIFT_X;
            const int x1 = min((xmax + RING_RADIUS - a.g.min_x) >> CELL_SHIFT_X, a.g.grid_w - 1);
            const int y1 = min((ymax + RING_RADIUS - a.g.min_y) >> CELL_SHIFT_Y, a.g.grid_h - 1);
            // run A: own grid row from the first active lane's record on; run B: the next grid row
            const uint32_t j0 = p0 + (uint32_t)first;
            const int len_a = (int)(__ldg(a.g.cell_start + (uint32_t)g * a.g.grid_w + x1 + 1) - j0);
            uint32_t start_b = 0;
            int len_b = 0;
            if (g + 1 <= y1) {
                start_b = __ldg(a.g.cell_start + (uint32_t)(g + 1) * a.g.grid_w + x0);
                len_b = (int)(__ldg(a.g.cell_start + (uint32_t)(g + 1) * a.g.grid_w + x1 + 1) - start_b);
            }
            const int total = len_a + len_b;
            const int need = active ? need_all : 33;                   // idle lanes never pass
            // candidate record v of the list; beyond it a pattern that is at least unlikely to pass (the
            // range is checked before a pair is queued)
            auto fetch = [&](int v) {
                return v < total ? __ldg(a.pre + (v < len_a ? j0 + (uint32_t)v : start_b + (uint32_t)(v - len_a)))
                                 : make_uint2(0x99999999u, 0x3c3c3c3cu);
            };
            uint2 next = fetch(lane);
            for (int v0 = 0; v0 < total; v0 += 32) {
                __syncwarp();
                s_pre[warp][lane] = next;
                __syncwarp();
                next = fetch(v0 + 32 + lane);                          // in flight while this chunk is tested
                uint32_t pass = 0;
                if (v0 == 0) {
                    // the warp's own records are candidates 0.. of this chunk (lane l's is l - first): collect
                    // the bit mask and keep only the records after one's own
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if ((int)__popc(c.matched(s_pre[warp][k].x, s_pre[warp][k].y)) >= need) pass |= 1u << k;
                    pass &= ~((2u << ((lane - first) & 31)) - 1u);
                } else {
                    int best = 0;
#pragma unroll
                    for (int k = 0; k < 32; k += 2)
                        best = __vimax3_s32(best, (int)__popc(c.matched(s_pre[warp][k].x, s_pre[warp][k].y)),
                                            (int)__popc(c.matched(s_pre[warp][k + 1].x, s_pre[warp][k + 1].y)));
                    // some candidate passed for these lanes (about one chunk in twenty): the warp finds out
                    // which, one such lane at a time -- its sets broadcast, every lane tests one candidate
                    uint32_t flagged = __ballot_sync(0xffffffffu, best >= need);
                    while (flagged) {
                        const int src = __ffs(flagged) - 1;
                        flagged &= flagged - 1;
                        Head32Sets o;
                        o.sa = __shfl_sync(0xffffffffu, c.sa, src);
                        o.sc = __shfl_sync(0xffffffffu, c.sc, src);
                        o.sg = __shfl_sync(0xffffffffu, c.sg, src);
                        o.st = __shfl_sync(0xffffffffu, c.st, src);
                        const uint2 b = s_pre[warp][lane];
                        const uint32_t hits = __ballot_sync(0xffffffffu, (int)__popc(o.matched(b.x, b.y)) >= need_all);
                        if (lane == src) pass = hits;
                    }
                }
                // pairs that passed go to the verify kernel (ring test, index windows, full-length compare)
                while (pass) {
                    const int k = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const int vk = v0 + k;
                    if (vk >= total) continue;
                    const uint32_t j = vk < len_a ? j0 + (uint32_t)vk : start_b + (uint32_t)(vk - len_a);
                    const uint32_t pos = atomicAdd(a.work_count, 1u);
                    if (pos < a.work_cap) a.work[pos] = make_uint2(p, j);
                }
            }
        }
    }
}

// ============================================================================
// per tile: exact test of the queued pairs, one pair per thread
// ============================================================================
struct XVerArgs {
    XGrid g;
    const uint64_t *packed;
    unsigned long long *tally;           // [n] grid order
    const uint2 *work;
    const uint32_t *work_count;
    uint32_t work_cap;
    int levels, len, e, hamming;
    uint32_t wlo, whi;
};

template <int W>
__global__ void __launch_bounds__(256)
exh_verify_kernel(XVerArgs a) {
    const uint32_t n_work = min(__ldg(a.work_count), a.work_cap);
    const int d2_max = c_x_d2[a.levels];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_work; i += gridDim.x * blockDim.x) {
        const uint2 pr = __ldg(a.work + i);
        const int4 me = __ldg(a.g.cell_wells + pr.x), w = __ldg(a.g.cell_wells + pr.y);
        const int dx = w.x - me.x, dy = w.y - me.y;
        const int d2 = dx * dx + dy * dy;
        if (!in_rings(d2, d2_max)) continue;
        // the scan window of the reference is not symmetric: each well of the pair has its own
        uint32_t lo, span, lo2, span2;
        index_window((uint32_t)me.z, a.wlo, a.whi, lo, span);
        index_window((uint32_t)w.z, a.wlo, a.whi, lo2, span2);
        const bool mine = (uint32_t)((uint32_t)w.z - lo) <= span;          // w is in me's window
        const bool theirs = (uint32_t)((uint32_t)me.z - lo2) <= span2;     // me is in w's window
        if (!mine && !theirs) continue;
        PSeq<W> cs, bs;
        load_packed<W>(a.packed + (size_t)(uint32_t)me.z * (W * PACK_STRIDE), cs);
        load_packed<W>(a.packed + (size_t)(uint32_t)w.z * (W * PACK_STRIDE), bs);
        if (is_duplicate<W>(cs, bs, a.len, a.e, a.hamming != 0)) {
            const unsigned long long one = 1ull << (XFIELD * ring_of(d2));
            if (mine) atomicAdd(a.tally + pr.x, one);
            if (theirs) atomicAdd(a.tally + pr.y, one);
        }
    }
}

template <int W>
static void launch_verify(const XVerArgs &a, unsigned blocks, cudaStream_t st) {
    exh_verify_kernel<W><<<blocks, 256, 0, st>>>(a);
}

// ============================================================================
// per tile: the sums of output_writer (count_well_duplicates.py:77-106)
// ============================================================================
__global__ void __launch_bounds__(256)
exh_finish_kernel(const unsigned long long *__restrict__ tally, const unsigned long long *__restrict__ ringlen,
                  uint32_t n, int L, int all_dups, unsigned long long *__restrict__ counters) {
    __shared__ uint32_t s_cnt[1 + 5 * 5];
    for (int i = threadIdx.x; i < 1 + 5 * 5; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    // per-thread sums over the grid-stride loop (ring sizes <= 4095 each, so 2^20 wells per thread fit 32 bits)
    uint32_t targets = 0, wells[5], dups[5], hit[5], acco[5], acci[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) wells[l] = dups[l] = hit[l] = acco[l] = acci[l] = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < n; p0 += 4 * stride) {
        unsigned long long tt[4], rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                       // eight loads in flight per thread
            const uint32_t p = p0 + u * stride;
            tt[u] = p < n ? __ldg(tally + p) : X_INVALID;
            rr[u] = p < n ? __ldg(ringlen + p) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned long long t = tt[u];
            if (t & X_INVALID) continue;
            const unsigned long long rl = rr[u];
            if (all_dups) t = rl;                           // e >= len: every ring well is a duplicate
            uint32_t hit_mask = 0;
#pragma unroll
            for (int l = 0; l < 5; ++l)
                if (l < L && ((t >> (XFIELD * l)) & XFIELD_MAX)) hit_mask |= 1u << l;
            targets++;
#pragma unroll
            for (int l = 0; l < 5; ++l) {
                if (l < L) {
                    wells[l] += (uint32_t)(rl >> (XFIELD * l)) & XFIELD_MAX;
                    dups[l] += (uint32_t)(t >> (XFIELD * l)) & XFIELD_MAX;
                    hit[l] += (hit_mask >> l) & 1u;
                    // AccO: a hit at this ring or further in; AccI: at this ring or further out
                    acco[l] += (hit_mask & ((2u << l) - 1u)) != 0;
                    acci[l] += (hit_mask >> l) != 0;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31;
    targets = __reduce_add_sync(0xffffffffu, targets);
    if (lane == 0 && targets) atomicAdd(&s_cnt[0], targets);
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        if (l < L) {
            const uint32_t v[5] = {__reduce_add_sync(0xffffffffu, wells[l]), __reduce_add_sync(0xffffffffu, dups[l]),
                                   __reduce_add_sync(0xffffffffu, hit[l]), __reduce_add_sync(0xffffffffu, acco[l]),
                                   __reduce_add_sync(0xffffffffu, acci[l])};
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (v[k]) atomicAdd(&s_cnt[1 + 5 * l + k], v[k]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1 + 5 * L; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(counters + i, (unsigned long long)s_cnt[i]);
}

static XGrid make_grid(wd_ctx *ctx) {
    XGrid g;
    g.cell_start = ctx->cell_start.as<uint32_t>();
    g.cell_wells = ctx->cell_wells.as<int4>();
    g.min_x = ctx->min_x; g.min_y = ctx->min_y; g.grid_w = ctx->grid_w; g.grid_h = ctx->grid_h;
    g.n = ctx->n_locs;
    return g;
}

// ring sizes of every well for (levels, window), kept until the next wd_locs_load
static int exhaustive_geometry(wd_ctx *ctx, int levels, uint32_t wlo, uint32_t whi) {
    if (!(ctx->x_geom_levels == levels && ctx->x_geom_wlo == wlo && ctx->x_geom_whi == whi)) {
        cudaStream_t st = ctx->stream;
        const uint32_t n = ctx->n_locs;
        WD_TRY(ctx->x_ringlen.reserve((size_t)n * 8));
        WD_TRY(ctx->x_flags.reserve(8));
        WD_CUDA(cudaMemsetAsync(ctx->x_flags.p, 0xff, 4, st));
        WD_CUDA(cudaMemsetAsync(ctx->x_flags.as<uint32_t>() + 1, 0, 4, st));
        XGeomArgs a;
        a.g = make_grid(ctx);
        a.ringlen = ctx->x_ringlen.as<unsigned long long>();
        a.first_empty = ctx->x_flags.as<uint32_t>();
        a.overflow = a.first_empty + 1;
        a.levels = levels; a.wlo = wlo; a.whi = whi;
        const unsigned blocks = (unsigned)std::min<size_t>(((size_t)n + XW * 32 - 1) / (XW * 32), (size_t)ctx->sm_count * 16);
        exh_geometry_kernel<<<blocks, XW * 32, 0, st>>>(a);
        ctx->launches++;
        WD_CUDA(cudaGetLastError());
        uint32_t fl[2];
        WD_CUDA(cudaMemcpyAsync(fl, ctx->x_flags.p, 8, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaStreamSynchronize(st));
        ctx->x_geom_first_empty = fl[0];
        ctx->x_geom_overflow = fl[1] != 0;
        ctx->x_geom_levels = levels; ctx->x_geom_wlo = wlo; ctx->x_geom_whi = whi;
    }
    if (ctx->x_geom_overflow)
        WD_FAIL(WD_E_ARG, "wd_count_exhaustive: more than %u wells in the grid cells around a well; the tile is denser than supported",
                XFIELD_MAX);
    if (ctx->x_geom_first_empty != UINT32_MAX)
        WD_FAIL(WD_E_RUNTIME, "Got no wells for cluster %u level %u", ctx->x_geom_first_empty / levels,
                ctx->x_geom_first_empty % levels);
    return WD_OK;
}

int count_exhaustive(wd_ctx *ctx, int slot, const int32_t *order, int seq_len, int levels, uint32_t wlo, uint32_t whi,
                     int e, int hamming, int64_t *tile_counters) {
    TileSlot &s = ctx->slots[slot];
    if (levels < 1 || levels > 5) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: levels must be 1..5 (MAX_DISTS defines 5 rings)");
    if (ctx->n_locs == 0) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: call wd_locs_load first");
    if (ctx->n_locs != s.n)
        WD_FAIL(WD_E_ASSERT, "wd_count_exhaustive: the .locs file holds %u wells, the tile %u", ctx->n_locs, s.n);
    if (s.n >= (1u << 29)) WD_FAIL(WD_E_ARG, "wd_count_exhaustive: at most 2^29 wells per tile");
    int words = 0;
    WD_TRY(pack_dense(ctx, slot, order, seq_len, &words));          // checks the plane order before anything is reported
    WD_TRY(exhaustive_geometry(ctx, levels, wlo, whi));
    cudaStream_t st = ctx->stream;
    const size_t width = 1 + 5 * (size_t)levels;
    WD_TRY(ctx->x_pre.reserve((size_t)s.n * 8));
    WD_TRY(ctx->x_tally.reserve((size_t)s.n * 8));
    WD_TRY(ctx->x_counts.reserve(width * 8 + 8));
    if (ctx->x_work_cap == 0) ctx->x_work_cap = std::max<size_t>((size_t)1 << 20, (size_t)s.n / 4);
    std::vector<unsigned long long> h(width + 1);
    for (int attempt = 0;; ++attempt) {
        WD_TRY(ctx->x_work.reserve(ctx->x_work_cap * 8));
        WD_CUDA(cudaMemsetAsync(ctx->x_counts.p, 0, width * 8 + 8, st));
        uint32_t *work_count = reinterpret_cast<uint32_t *>(ctx->x_counts.as<unsigned long long>() + width);
        exh_prefix_kernel<<<(s.n + 255) / 256, 256, 0, st>>>(ctx->cell_wells.as<int4>(), ctx->x_packed.as<uint64_t>(), words,
                                                              s.n, ctx->x_pre.as<uint2>(), ctx->x_tally.as<unsigned long long>());
        ctx->launches++;
        // e < 0: no pair is a duplicate; e >= len: every pair is one -- neither needs a sequence
        if (e >= 0 && e < seq_len) {
            XCmpArgs a;
            a.g = make_grid(ctx);
            a.pre = ctx->x_pre.as<uint2>();
            a.packed = ctx->x_packed.as<uint64_t>();
            a.work = ctx->x_work.as<uint2>();
            a.work_count = work_count;
            a.work_cap = (uint32_t)std::min<size_t>(ctx->x_work_cap, 0xffffffffu);
            a.words = words; a.e = e;
            a.k = head32_k(e, hamming != 0);
            // one warp per 32 centres; many short CTAs per SM so that the tail is short
            const unsigned blocks = (unsigned)std::min<size_t>(((size_t)s.n + XW * 32 - 1) / (XW * 32), (size_t)ctx->sm_count * 64);
            exh_compare_kernel<<<blocks, XW * 32, 0, st>>>(a);
            ctx->launches++;
            XVerArgs v;
            v.g = a.g; v.packed = a.packed; v.tally = ctx->x_tally.as<unsigned long long>();
            v.work = a.work; v.work_count = work_count; v.work_cap = a.work_cap;
            v.levels = levels; v.len = seq_len; v.e = e; v.hamming = hamming; v.wlo = wlo; v.whi = whi;
            const unsigned vblocks = (unsigned)ctx->sm_count * 2;
            switch (words) {
                case 1: launch_verify<1>(v, vblocks, st); break;
                case 2: launch_verify<2>(v, vblocks, st); break;
                case 4: launch_verify<4>(v, vblocks, st); break;
                case 8: launch_verify<8>(v, vblocks, st); break;
                default: launch_verify<16>(v, vblocks, st); break;
            }
            ctx->launches++;
        }
        exh_finish_kernel<<<(unsigned)std::min<size_t>(((size_t)s.n + 255) / 256, (size_t)ctx->sm_count * 8), 256, 0, st>>>(
            ctx->x_tally.as<unsigned long long>(), ctx->x_ringlen.as<unsigned long long>(), s.n, levels,
            (e >= 0 && e >= seq_len) ? 1 : 0, ctx->x_counts.as<unsigned long long>());
        ctx->launches++;
        WD_CUDA(cudaGetLastError());
        WD_CUDA(cudaMemcpyAsync(h.data(), ctx->x_counts.p, (width + 1) * 8, cudaMemcpyDeviceToHost, st));
        WD_CUDA(cudaStreamSynchronize(st));
        const size_t queued = (size_t)(h[width] & 0xffffffffull);
        if (queued <= ctx->x_work_cap) break;
        // more pairs passed the 32-symbol test than the queue holds (low-complexity reads): grow it and redo the tile
        if (attempt > 0) WD_FAIL(WD_E_CAPACITY, "wd_count_exhaustive: pair queue overflow (%zu pairs)", queued);
        ctx->x_work_cap = queued + queued / 8 + 1024;
    }
    for (size_t i = 0; i < width; ++i) tile_counters[i] = (int64_t)h[i];
    return WD_OK;
}

}  // namespace wd
