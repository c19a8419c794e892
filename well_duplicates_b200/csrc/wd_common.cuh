// Internal declarations shared by the translation units of libwelldup.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/welldup.h"

namespace wd {

// ---- error plumbing --------------------------------------------------------
void set_error(const char *fmt, ...);
#define WD_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            wd::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                          cudaGetErrorString(e__));                                        \
            return WD_E_CUDA;                                                              \
        }                                                                                  \
    } while (0)
#define WD_TRY(call)                    \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != WD_OK) return rc__; \
    } while (0)
#define WD_FAIL(code, ...)           \
    do {                             \
        wd::set_error(__VA_ARGS__);  \
        return (code);               \
    } while (0)

// ---- device buffer that only grows -------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);   // keeps contents only when no reallocation happens
    void release();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// ---- stage-1 grid ------------------------------------------------------------------
// Cells are 32 px wide and 128 px tall.  128 >= the largest ring radius (102 px), so three
// grid rows cover every ring; inside a grid row the cells are contiguous in memory, so the
// wells with |dx| <= 102 are ONE run of records (cells (cx-102)>>5 .. (cx+102)>>5, 7-8 narrow
// cells, ~250 px instead of the 384 px three square cells would span).
constexpr int CELL_SHIFT_X = 5;
constexpr int CELL_SHIFT_Y = 7;
constexpr int RING_RADIUS = 102;                    // MAX_DISTS[-1], prepare_cluster_indexes.py:19

// ---- one tile's planes in HBM ---------------------------------------------------
struct TileDesc {                 // what the kernels see (array in device memory)
    const uint8_t *planes;        // n_planes x stride bytes
    unsigned long long stride;    // bytes between planes (multiple of 256)
    const uint8_t *filter;        // n bytes, bit0 = PF
    const uint64_t *pfmask;       // ceil(n/64) words: PF bit per well      (K3)
    const uint32_t *pfrank;       // ceil(n/64) words: #PF wells before block (K3)
    const uint8_t *kind;          // n_planes bytes: WD_PLANE_*
    uint32_t n;                   // wells on the tile
    uint32_t flags;               // bit0: pfmask / pfrank are valid; bit1: excl_expect is to be checked
    uint32_t excl_expect;         // cluster count of the tile's excluded CBCL blocks: must equal the PF total (cbcl_read.py:130-131)
    uint32_t reserved;
    // host-mapped tiles only: the planes of the first compared positions are copied to HBM by DMA
    // (head[j] = plane of position j); head_delta = head - planes (mod 2^64), so planes + head_delta
    // + j * head_stride addresses them
    unsigned long long head_delta;
    unsigned long long head_stride;
};

struct TileSlot {
    uint32_t n = 0;
    int n_planes = 0;
    size_t stride = 0;
    DevBuf planes, filter, pfmask, pfrank, kind_dev;
    uint8_t *head_ptr = nullptr;       // host-mapped: this tile's head planes inside ctx->head for the count being issued
    size_t head_stride = 0;
    const uint8_t *mapped = nullptr;   // planes left in pinned host memory (wd_tile_map_host): device view of it
    const uint8_t *mapped_host = nullptr;     // ... and the host view (source of DMA copies)
    const uint8_t *mapped_filter = nullptr;   // same for the filter bytes (optional)
    const uint8_t *mapped_filter_host = nullptr;
    std::vector<uint8_t> kind;
    std::vector<uint32_t> n_block;
    bool filter_set = false;
    bool rank_valid = false;
    bool kind_dirty = true;
    bool has_excl = false;
};

// ---- target list on the device ------------------------------------------------------
// Slots are the wells of all targets laid out target after target: the centre
// first, then the ring wells sorted by well index (so that consecutive lanes
// read neighbouring bytes of a plane); slot_level says which ring a slot
// belongs to, slot_csr where it sat in the caller's CSR (reference order).
struct TargetList {
    uint32_t t = 0;
    int levels = 0;
    uint32_t n_slots = 0;
    uint32_t max_well = 0;
    DevBuf tgt_off;      // u32 [t+1]
    DevBuf slot_well;    // u32 [n_slots]
    DevBuf slot_level;   // u8  [n_slots]
    DevBuf slot_csr;     // u32 [n_slots]  (centre: UINT32_MAX)
    DevBuf level_len;    // u32 [t*levels] wells per ring (LENGTH of the reference)
    DevBuf visit;        // u32 [t] targets in ascending order of their centre well: the fused kernel walks
                         //         them in this order, so that a CTA's targets sit on neighbouring rows of a
                         //         plane (few pages / DRAM rows per CTA); results are stored by target ordinal
    std::vector<uint32_t> h_idx;        // host copy of the caller's idx[] (duplicate-pair log)
    std::vector<uint32_t> h_slot_csr;   // ... and of slot_csr: a logged slot -> position in the caller's list
    bool has_empty_ring = false;        // some target has a ring without wells: an AssertionError of the reference
                                        // (count_well_duplicates.py:249) -- but only once such a target's centre
                                        // passes the filter of a tile that is counted
};

// knobs of wd_count (wd_set_tuning); 0 / negative = the library's choice
struct Tuning {
    int step0 = 0, step1 = 0;      // cycles read per round of the fused kernel (first, later), 1..8
    int centre_chunk = 0;          // centre cycles decoded per warp-wide load: 8, 16 or 32
    int head_planes = -1;          // host-mapped tiles: compared positions whose planes go to HBM by DMA, 0..8
    int head_groups = 0;           // ... in how many tile groups, pipelined against the counting kernels
    int visit_order = -1;          // 0: targets in list order, 1 / -1: in ascending order of their centre well
    int targets_per_cta = 0;       // fused kernel: targets per CTA (8..256), 0 = chosen per launch
    int ctas_per_sm = 0;           // fused kernel: resident CTAs per SM capped by reserving shared memory, 0 = no cap
};

}  // namespace wd

struct wd_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // DMA of head planes, overlapped with the counting kernels
    std::vector<cudaEvent_t> copy_events;
    int sm_count = 148;
    uint64_t launches = 0;

    // stage 1
    uint32_t n_locs = 0;
    wd::DevBuf xy, px, py, bbox, cell_start, cell_cursor, cell_wells, scan_tmp;
    int grid_w = 0, grid_h = 0, min_x = 0, min_y = 0;
    wd::DevBuf q_centres, q_counts, q_offsets, q_idx, q_tmp, q_flag;
    uint32_t q_t = 0;
    int q_levels = 0;
    uint64_t q_total = 0;

    // stage 2/3
    std::vector<wd::TileSlot> slots;
    wd::TargetList targets;
    wd::DevBuf descs, order_dev, packed, per_target, counters, publish, dup_rows, dup_count;
    wd::DevBuf gs_idx, gs_packed, gs_codes;
    // exhaustive mode (wd_exhaustive.cu)
    wd::DevBuf x_packed, x_counts, x_pre, x_ringlen, x_flags, x_tally, x_work;
    size_t x_work_cap = 0;            // pairs the queue between exh_compare_kernel and exh_verify_kernel holds
    int x_geom_levels = 0;            // ring sizes in x_ringlen are valid for (levels, window); 0 = none
    uint32_t x_geom_wlo = 0, x_geom_whi = 0, x_geom_first_empty = 0;
    bool x_geom_overflow = false;
    // last count
    int last_tiles = 0, last_levels = 0, last_first_slot = 0;
    uint32_t last_t = 0;
    bool last_per_target = false;
    size_t publish_n = 0;
    std::vector<unsigned long long> order_off;   // plane order / kinds / tile descriptors as last uploaded
    std::vector<uint8_t> order_kind, descs_host;
    std::vector<int32_t> publish_map;   // tile_row ++ lane_row as last uploaded behind the publish buffer
    size_t dup_cap = 0;               // rows the duplicate-pair log of the last count has room for (0: not logged)
    size_t dup_cap_wanted = 0;        // ... and what a repeat of it after an overflow asks for
    std::vector<int32_t> dup_raw;     // rows of the last logged count once fetched (wd_dup_pairs is called twice: size, rows)
    bool dup_raw_valid = false;
    uint64_t last_h2d_bytes = 0;      // bytes the last wd_count copied to HBM itself (head planes of host-mapped tiles)
    wd::Tuning tuning;
    // arguments of the last count, kept so that wd_dup_pairs can grow the log and run it again
    std::vector<int32_t> last_order;
    int last_e = 0, last_hamming = 0, last_mode = 0, last_seq_len = 0;
    bool last_all_bcl = true;
    wd::DevBuf head;                  // host-mapped tiles: [tile][head plane][head stride] of the last count
    cudaEvent_t dma_ev0 = nullptr, dma_ev1 = nullptr;   // around the head-plane copies of a count (timing)
    bool dma_pending = false;
    uint64_t dma_bytes_timed = 0;
    double dma_gbps = 0.0;            // rate of the head-plane copies of an earlier count (0: not measured yet)
    int last_n_head = 0;
    wd::DevBuf trace, trace_counts;   // wd_count_trace_sectors
    wd::DevBuf dup_codes;             // wd_dup_pairs_seqs
    wd::DevBuf rank_jobs;             // K3: job descriptors of a batch of tiles
    // multi-GPU (wd_comm.cc)
    void *comm = nullptr;             // ncclComm_t
    int comm_rank = 0, comm_ranks = 1;
    cudaStream_t comm_stream = nullptr;          // the all-reduce of one step runs under the kernels of the next
    cudaEvent_t pub_ready = nullptr, comm_done[2] = {nullptr, nullptr};
    bool comm_pending[2] = {false, false};       // an all-reduce of publish buffer b is in flight on comm_stream
    int publish_cur = 0;                         // publish buffer of the last wd_publish_counters (two alternate)
};

namespace wd {

// stage 1 (wd_stage1.cu)
int locs_load(wd_ctx *ctx, const float *xy, uint32_t n);
int ring_query(wd_ctx *ctx, const uint32_t *centres, uint32_t t, int levels, uint32_t wlo,
               uint32_t whi, uint32_t *level_offsets, uint32_t *idx, size_t idx_cap,
               uint64_t *n_idx, uint32_t *first_empty);

// stage 2/3 (wd_stage23.cu)
int filter_rank(wd_ctx *ctx, const int *slot_ids, int n);
int filter_offsets(wd_ctx *ctx, int slot, int32_t *offsets, uint32_t *passing);
int get_seqs(wd_ctx *ctx, int slot, const int64_t *indices, uint32_t n_idx, const int32_t *order,
             int seq_len, uint8_t *codes, uint8_t *pf);
int count_async(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e,
                int hamming, int mode, int want_per_target);
int count_trace(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *order, int seq_len, int e,
                int hamming, uint32_t *sectors, uint32_t *lines);
int dup_rows_fetch(wd_ctx *ctx, std::vector<int32_t> &raw);
int dup_seqs(wd_ctx *ctx, const std::vector<int32_t> &raw, std::vector<uint8_t> &codes);
int comm_destroy(wd_ctx *ctx);
int publish_counters(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row, int n_tiles,
                     int n_rows_total, bool keep);
int count_exhaustive(wd_ctx *ctx, int slot, const int32_t *order, int seq_len, int levels,
                     uint32_t wlo, uint32_t whi, int e, int hamming, int64_t *tile_counters);
int upload_descs(wd_ctx *ctx, int first_slot, int n_tiles);
int pack_dense(wd_ctx *ctx, int slot, const int32_t *order, int seq_len, int *words_out);
}  // namespace wd
