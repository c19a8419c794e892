/*
 * welldup.h -- C ABI of libwelldup.so, the B200 (sm_100a) implementation of the
 * well_duplicates hot path.
 *
 * The reference (EdinburghGenomics/well_duplicates) is pure Python and has no
 * FFI; the seams this library sits behind are its two CLIs and its Python
 * reader API.  Each entry point below names the reference code it replaces
 * (file:line under the reference tree).  The ctypes binding that a maintainer
 * of the reference would add is shown in INTEGRATION.md and shipped as
 * well_duplicates_b200/_lib.py.
 *
 * Conventions
 *  - every function returns an int status: WD_OK or a negative WD_E_* class
 *    that tells the Python side which exception the reference would have
 *    raised; the text is available from wd_last_error() (thread local);
 *  - nothing throws across the boundary;
 *  - host pointers are borrowed: pageable memory is consumed before the call
 *    returns; memory obtained from wd_host_alloc() (pinned) is read
 *    asynchronously and must stay unchanged until the next wd_sync() or
 *    result-returning call (wd_count, wd_get_seqs, ...) on that context;
 *  - all device memory belongs to the wd_ctx and is released by wd_destroy();
 *  - output arrays are allocated by the caller;
 *  - one wd_ctx per GPU, not thread-safe; work is asynchronous on the
 *    context's stream until a result-returning call or wd_sync().
 *
 * There is no CPU fallback: wd_create() fails when no CUDA device is usable.
 */
#ifndef WELLDUP_H
#define WELLDUP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WD_ABI_VERSION 2

#if defined(__GNUC__)
#define WD_API __attribute__((visibility("default")))
#else
#define WD_API
#endif

#define WD_OK 0
#define WD_E_INDEX (-1)    /* IndexError: well index out of range (bcl_direct_reader.py:186-192) */
#define WD_E_RUNTIME (-2)  /* RuntimeError: a target has an empty ring (prepare_cluster_indexes.py:70-76) */
#define WD_E_ASSERT (-3)   /* AssertionError: header / size mismatch (bcl_direct_reader.py:338, :236) */
#define WD_E_CUDA (-4)     /* CUDA runtime failure */
#define WD_E_ARG (-5)      /* ValueError: argument outside what the library supports */
#define WD_E_CAPACITY (-6) /* caller's output array too small; required size is reported */
#define WD_E_NOENT (-7)    /* FileNotFoundError: the reader then tries the CBCL file (bcl_direct_reader.py:209-216) */
#define WD_E_IO (-8)       /* OSError: open / read failed */
#define WD_E_EOF (-9)      /* EOFError: compressed data end before the end-of-stream marker (gzip.open().read()) */
#define WD_E_DATA (-10)    /* gzip.BadGzipFile / zlib.error: not gzip, corrupt deflate data, CRC or length mismatch */

#define WD_MAX_LEVELS 15   /* rings per target (the reference's MAX_DISTS gives 5) */
#define WD_MAX_SEQ_LEN 1024 /* compared symbols per well (sum of all --cycles ranges) */

/* plane kinds for wd_tile_put_* */
#define WD_PLANE_EMPTY 0
#define WD_PLANE_BCL 1       /* 1 byte / well */
#define WD_PLANE_CBCL 2      /* 4 bits / well, every well stored */
#define WD_PLANE_CBCL_EXCL 3 /* 4 bits / well, pass-filter wells only */

typedef struct wd_ctx wd_ctx;

/* ---- context -------------------------------------------------------------- */
WD_API int wd_abi_version(void);
WD_API const char *wd_last_error(void);
WD_API int wd_create(int device, wd_ctx **out);
WD_API int wd_destroy(wd_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of
 * the context's own; pass NULL to return to the private stream. */
WD_API int wd_set_stream(wd_ctx *ctx, void *cuda_stream);
WD_API int wd_sync(wd_ctx *ctx);
/* Pinned host memory, so that wd_tile_put_* / wd_locs_load copy by DMA without staging. */
WD_API int wd_host_alloc(size_t bytes, void **out);
WD_API int wd_host_free(void *p);
/* Page-lock (and map) memory the caller already owns -- then it can be passed to wd_tile_map_host and is copied
 * by DMA like memory from wd_host_alloc().  Unregister before freeing it. */
WD_API int wd_host_register(void *p, size_t bytes);
WD_API int wd_host_unregister(void *p);
/* DRAM->L2 fill granularity hint for the current device (32, 64 or 128 bytes;
 * cudaLimitMaxL2FetchGranularity).  The scattered plane gathers use a few bytes
 * per 32-byte sector, so wd_create() asks for 32; *previous receives the old value. */
WD_API int wd_set_l2_fetch_granularity(wd_ctx *ctx, int bytes, int *previous);
/* Bytes the last wd_count / wd_count_async copied from host-mapped tiles to HBM by DMA (the planes
 * of the first compared positions, see wd_tile_map_host); 0 for staged tiles. */
WD_API int wd_last_count_h2d_bytes(wd_ctx *ctx, uint64_t *out);
/* Host-mapped tiles: how many leading planes the last wd_count copied by DMA (1 unless wd_set_tuning says
 * otherwise) and the measured rate of an earlier count's copies in GB/s (0: not measured yet). */
WD_API int wd_last_count_staging(wd_ctx *ctx, int *head_planes, double *dma_gb_per_s);
/* Number of kernel launches issued by this context so far (bench.py gpu_launches). */
WD_API int wd_launch_count(wd_ctx *ctx, uint64_t *out);
/* Knobs of wd_count for measurement sweeps (profiles/): 0 (or -1 where noted) leaves the library's choice.
 * NULL restores every default.  The library reads no environment variables. */
typedef struct wd_tuning {
    int32_t step0, step1;  /* cycles the fused kernel reads per round (first round, later rounds), 1..8 */
    int32_t centre_chunk;  /* 8, 16 or 32: the centre is read ahead to a multiple of this many cycles (default: exactly as far as needed) */
    int32_t head_planes;   /* host-mapped tiles: compared positions whose planes are copied to HBM by DMA, 0..8; -1 default */
    int32_t head_groups;   /* ... in how many tile groups, pipelined against the counting kernels */
    int32_t visit_order;   /* 0: targets in list order; 1 or -1: in ascending order of their centre well */
    int32_t targets_per_cta; /* fused kernel: targets one CTA's 8 warps share, a multiple of 8 in 8..256; 0 = the library's choice */
    int32_t ctas_per_sm;     /* fused kernel: at most this many CTAs resident per SM (1..8), by reserving shared memory; 0 = no limit */
    int32_t reserved[8];
} wd_tuning;
WD_API int wd_set_tuning(wd_ctx *ctx, const wd_tuning *tuning);

/* ---- stage 1: neighbourhood construction ----------------------------------
 * Replaces yield_coords + get_indexes of prepare_cluster_indexes.py
 * (:99-116, :38-78): K0 converts the .locs floats with the reference's float64
 * formula, K1 bins the wells into a uniform grid of 128-px cells, K2 searches
 * the 3x3 cells around each centre. */
WD_API int wd_locs_load(wd_ctx *ctx, const float *xy /* n x 2 */, uint32_t n);
/* parity hook: the integer pixel coordinates K0 produced (prepare_cluster_indexes.py:110-112) */
WD_API int wd_locs_pixels(wd_ctx *ctx, int32_t *x /* n */, int32_t *y /* n */);
/* Rings 1..levels around each centre, using the index window
 * [c - window_lo, c + window_hi] (20000 / 20001 in the reference, :43,:52-67)
 * and the distance bins of MAX_DISTS (:19,:61-63); indices ascend inside a
 * level.  level_offsets has t*levels+1 entries; idx receives
 * level_offsets[t*levels] entries.  If idx_cap is too small the call returns
 * WD_E_CAPACITY with *n_idx set to the size needed.  If some ring is empty the
 * call returns WD_E_RUNTIME and *first_empty = target*levels + level of the
 * first such ring in sample order (UINT32_MAX otherwise). */
WD_API int wd_ring_query(wd_ctx *ctx, const uint32_t *centres, uint32_t t, int levels,
                  uint32_t window_lo, uint32_t window_hi,
                  uint32_t *level_offsets, uint32_t *idx, size_t idx_cap,
                  uint64_t *n_idx, uint32_t *first_empty);

/* ---- target list ------------------------------------------------------------
 * The CSR form of what target.py:load_targets returns (:6-40): centres[t],
 * level_offsets[t*levels+1] into idx[], for rings 1..levels. */
WD_API int wd_targets_load(wd_ctx *ctx, const uint32_t *centres, const uint32_t *level_offsets,
                    const uint32_t *idx, uint32_t t, int levels);

/* ---- stage 2: tile staging ----------------------------------------------------
 * A tile slot holds the gunzipped planes of one tile in HBM.  Replaces the
 * per-cycle slurp in Tile.get_seqs (bcl_direct_reader.py:200-216, :333-345,
 * :300-301); gunzip itself stays on the host. */
WD_API int wd_tile_begin(wd_ctx *ctx, int tile_slot, uint32_t n_clusters, int n_planes);
/* body of the .filter file (header stripped): bit0 = pass filter (:222-253) */
WD_API int wd_tile_put_filter(wd_ctx *ctx, int tile_slot, const uint8_t *bytes, uint32_t n);
/* body of a gunzipped .bcl (4-byte count stripped); n must equal n_clusters (:333-338) */
WD_API int wd_tile_put_bcl(wd_ctx *ctx, int tile_slot, int plane, const uint8_t *bytes, uint32_t n);
/* one inflated CBCL tile block (:300-301); n_block = cluster count in the
 * tile record (PF count when excluded != 0) */
WD_API int wd_tile_put_cbcl(wd_ctx *ctx, int tile_slot, int plane, const uint8_t *nibbles,
                     uint32_t usize, uint32_t n_block, int excluded);
/* Zero-copy staging, instead of wd_tile_begin + wd_tile_put_bcl/_cbcl: the
 * tile's inflated planes stay where the host put them -- planes[p * stride_bytes]
 * in page-locked memory from wd_host_alloc().  wd_count then copies only the
 * plane of the first compared position to HBM (DMA, overlapped tile
 * group by tile group with the kernels) and the counting kernel pulls the
 * 32-byte sectors it needs of the later planes straight across PCIe -- a few
 * per cent of a sampled tile instead of the whole 215 MB the per-cycle slurp
 * of bcl_direct_reader.py:333-345 moves.  kinds[p] is a WD_PLANE_* (NULL = all
 * BCL), n_block[p] the cluster count of a CBCL block (NULL = n_clusters).  The
 * memory must stay unchanged until the results of the last wd_count that uses
 * the slot have been fetched.  filter (n_clusters bytes, body of the .filter
 * file) may be mapped the same way, or NULL: then wd_tile_put_filter supplies it. */
WD_API int wd_tile_map_host(wd_ctx *ctx, int tile_slot, uint32_t n_clusters, int n_planes,
                     const uint8_t *planes, size_t stride_bytes, const uint8_t *kinds,
                     const uint32_t *n_block, const uint8_t *filter);
/* K3: filter byte -> rank among PF wells or -1 (Tile._get_filter_offsets, :222-253) */
WD_API int wd_filter_offsets(wd_ctx *ctx, int tile_slot, int32_t *offsets /* n_clusters */,
                      uint32_t *passing);
/* K3+K4/K5: Tile.get_seqs (:158-220) for arbitrary wells.  plane_order[p] is
 * the plane that supplies sequence position p.  codes[i*seq_len + p] is
 * 0..3 = A,C,G,T or 4 = N; pf[i] is the QUAL_FLAG.  Out-of-range or negative
 * indices give WD_E_INDEX. */
WD_API int wd_get_seqs(wd_ctx *ctx, int tile_slot, const int64_t *indices, uint32_t n_idx,
                const int32_t *plane_order, int seq_len, uint8_t *codes, uint8_t *pf);

/* ---- stage 3: compare + count -------------------------------------------------
 * Replaces the target / level / well loops of main() and the integer part of
 * output_writer (count_well_duplicates.py:228-265, :65-106) for tile slots
 * first_slot .. first_slot+n_tiles-1 against the loaded target list.
 *   per_target   [n_tiles][t][1+2*levels] int32: valid, then (dups, wells) per
 *                level -- the reference's lane_dupl structure; may be NULL.
 *   tile_counters[n_tiles][1+5*levels] int64: Targets, then per level
 *                Wells, Dups, Hit, AccO, AccI.
 * edit_distance / hamming are -e / --hamming (:200, :257).
 * mode: WD_MODE_FUSED = fused gather+compare kernel, WD_MODE_TWO_PASS = K4/K5 packed
 * words in HBM, then K6 (also logs), WD_MODE_FUSED_LOG = the fused kernel, and every
 * duplicate pair is logged for wd_dup_pairs (count_well_duplicates.py:258-262).  All
 * three give identical counters.  A target with an empty ring whose centre passes the
 * filter of a counted tile gives WD_E_ASSERT (:249) when the results are fetched. */
#define WD_MODE_FUSED 0
#define WD_MODE_TWO_PASS 1
#define WD_MODE_FUSED_LOG 2
WD_API int wd_count(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order, int seq_len,
             int edit_distance, int hamming, int mode,
             int32_t *per_target, int64_t *tile_counters);
/* Same, but only enqueues the kernels: results stay on the device (see
 * wd_counters_devptr) until wd_count_fetch. */
WD_API int wd_count_async(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order,
                   int seq_len, int edit_distance, int hamming, int mode, int want_per_target);
WD_API int wd_count_fetch(wd_ctx *ctx, int32_t *per_target, int64_t *tile_counters);
/* Duplicate pairs of the last wd_count in mode 1 or 2 (count_well_duplicates.py:258-262):
 * rows of (tile of the batch, target ordinal, well, distance), in reference log order
 * (tile; then target, level, well as listed).  *n_rows is always set; WD_E_CAPACITY when
 * cap is too small.  Every pair is delivered however many there are: if the device-side
 * log was too small the library grows it and repeats the count (the tile slots must still
 * hold their planes, i.e. call this before staging the next batch). */
WD_API int wd_dup_pairs(wd_ctx *ctx, int32_t *rows /* cap x 4 */, size_t cap, uint64_t *n_rows);
/* Same, plus the two sequences of every pair as the reference prints them (:260-261):
 * codes[row][0] = centre, codes[row][1] = ring well, seq_len bytes each, 0..3 = ACGT, 4 = N. */
WD_API int wd_dup_pairs_seqs(wd_ctx *ctx, int32_t *rows /* cap x 4 */, uint8_t *codes /* cap x 2 x seq_len */,
                      size_t cap, uint64_t *n_rows);
/* Measurement hook (bench.py roofline): runs the fused kernel of wd_count on the same
 * arguments in a build that records every plane read, and returns per tile and compared
 * position the number of distinct 32-byte sectors (and distinct 128-byte lines) of that
 * position's plane the kernel asked for -- with its early exits, i.e. what the algorithm
 * needs, not what a full read would move.  <= 64 compared symbols, <= 5 levels. */
WD_API int wd_count_trace_sectors(wd_ctx *ctx, int first_slot, int n_tiles, const int32_t *plane_order, int seq_len,
                           int edit_distance, int hamming, uint32_t *sectors /* n_tiles x seq_len */,
                           uint32_t *lines /* n_tiles x seq_len */);

/* ---- multi-GPU ------------------------------------------------------------------
 * One process per GPU, tiles sharded over the ranks; the counter rows are combined by
 * ONE ncclAllReduce(int64, sum) over NVLink.  Replaces one process per lane plus the
 * "tail" concatenation of Snakefile.count_dups:146-160.
 * K7: place this rank's tile counters into a zero-initialised
 * [n_rows_total][1+5*levels] int64 device array (tile_row[i] = global row of
 * local tile i, lane_row[i] = row that accumulates its lane total), ready for
 * wd_allreduce_i64 (or any other all-reduce on *devptr). */
WD_API int wd_publish_counters(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row,
                        int n_tiles, int n_rows_total, void **devptr, size_t *n_int64);
/* Same, into the array of the previous wd_publish_counters (not zeroed again): a rank that counts its
 * tiles in several batches publishes each batch, then reduces once. */
WD_API int wd_publish_add(wd_ctx *ctx, const int32_t *tile_row, const int32_t *lane_row,
                   int n_tiles, int n_rows_total, void **devptr, size_t *n_int64);
WD_API int wd_counters_devptr(wd_ctx *ctx, void **devptr, size_t *n_int64);
/* NCCL communicator of this context (the library loads libnccl.so.2 on first use; no torch).
 * Rank 0 calls wd_comm_unique_id and hands the 128 bytes to the other ranks by any means
 * (a file, MPI, torch.distributed's store); every rank then calls wd_comm_init. */
#define WD_COMM_ID_BYTES 128
WD_API int wd_comm_unique_id(void *id128);
WD_API int wd_comm_init(wd_ctx *ctx, const void *id128, int rank, int nranks);
WD_API int wd_comm_destroy(wd_ctx *ctx);
/* In-place sum over all ranks of n int64 at device pointer buf, enqueued on the context's stream.
 * buf == NULL: the array of the last wd_publish_counters -- reduced on a communication stream of the
 * library's own, behind the kernel that filled it, so that the next wd_count (whose rows go to a
 * second array) runs while the collective is in flight. */
WD_API int wd_allreduce_i64(wd_ctx *ctx, void *buf, size_t n);
/* Make the context's stream wait for the all-reduces issued so far (before timing or reusing results). */
WD_API int wd_comm_join(wd_ctx *ctx);
/* The (all-reduced) rows of the last wd_publish_counters, copied to the host (synchronises). */
WD_API int wd_published_fetch(wd_ctx *ctx, int64_t *rows /* n_int64 */, size_t n_int64);

/* ---- exhaustive mode ---------------------------------------------------------------
 * Every well of the tile is a target out to `levels` rings (BASELINE config 3):
 * neighbours come straight from the stage-1 grid (no materialised target list),
 * every well is packed once by a dense pass over the planes. */
WD_API int wd_count_exhaustive(wd_ctx *ctx, int tile_slot, const int32_t *plane_order, int seq_len,
                        int levels, uint32_t window_lo, uint32_t window_hi,
                        int edit_distance, int hamming, int64_t *tile_counters);

/* ---- host staging: gunzip -----------------------------------------------------------
 * Host-only (no GPU needed).  Replaces gzip.open(file).read() of the per-cycle
 * .bcl.gz slurp (bcl_direct_reader.py:207-208, :333-345) and the seek + gzip
 * member read of a CBCL tile block (:292-301): each job reads `size` bytes at
 * `offset` of `path` (size 0 = to the end of the file) -- or takes src[0, size)
 * when path is NULL -- and inflates every gzip member found there to dst, which
 * may point into page-locked memory from wd_host_alloc() so that wd_tile_map_host
 * can hand the planes to the kernels without another copy.  Jobs are independent
 * and are spread over `threads` native threads (0 = one per hardware thread).
 * gzip.open() semantics: CRC-32 and length of every member are checked, members
 * may follow each other, zero padding after a member is skipped, an empty input
 * gives no bytes.  Per-job status: WD_OK, WD_E_NOENT, WD_E_IO, WD_E_EOF,
 * WD_E_DATA, WD_E_CAPACITY (more than dst_cap bytes), WD_E_ARG; the call returns
 * the status of the first failed job (its message in wd_last_error()). */
typedef struct wd_inflate_job {
    const char *path;      /* file to read, or NULL */
    const uint8_t *src;    /* compressed bytes when path is NULL */
    uint64_t offset;       /* first byte of the gzip member(s) in the file */
    uint64_t size;         /* compressed bytes (0 with a path = to the end of the file) */
    uint8_t *dst;          /* where the inflated bytes go */
    uint64_t dst_cap;      /* room at dst */
    uint64_t out_len;      /* [out] bytes written to dst */
    int32_t status;        /* [out] WD_OK or WD_E_* */
    char message[220];     /* [out] text of the failure */
} wd_inflate_job;
WD_API int wd_inflate_batch(wd_inflate_job *jobs, size_t n_jobs, int threads);
/* one buffer, calling thread */
WD_API int wd_gunzip(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);
/* CRC-32 of RFC 1952 (zlib.crc32): running value in, running value out */
WD_API uint32_t wd_crc32(uint32_t crc, const uint8_t *data, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* WELLDUP_H */
