/* CPU restatement of the reference's hot path in plain C -- TEST INFRASTRUCTURE
 * (checker for tests/ and smoke(), and the timed "port" CPU baseline of
 * bench.py).  Never linked into, imported by or called from the product
 * (well_duplicates_b200/).  Each function cites the reference lines it follows;
 * oracle/ref_port.py holds the same logic in Python and both are pinned to the
 * golden outputs of the unmodified reference (tests/test_oracle_golden.py,
 * tests/test_oracle_c.py).
 *
 * PARITY UNPINNED at one boundary: Levenshtein.distance / .hamming
 * (count_well_duplicates.py:200,252) come from the third-party
 * python-Levenshtein C extension, unpinned and absent from the reference tree;
 * orc_levenshtein restates the published unit-cost definition.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int MAX_DISTS[6] = {1, 22, 42, 62, 82, 102}; /* prepare_cluster_indexes.py:19 */
#define MAX_SEARCH_AREA 20000                             /* prepare_cluster_indexes.py:43 */

/* prepare_cluster_indexes.py:110-112: int(t * 10.0 + 1000.5), float64, truncation */
void orc_locs_to_pixels(const float *xy, uint32_t n, int32_t *x, int32_t *y) {
    for (uint32_t i = 0; i < n; ++i) {
        volatile double vx = (double)xy[2 * i] * 10.0;
        volatile double vy = (double)xy[2 * i + 1] * 10.0;
        x[i] = (int32_t)(vx + 1000.5);
        y[i] = (int32_t)(vy + 1000.5);
    }
}

/* get_indexes, prepare_cluster_indexes.py:38-78: scan records
 * max(0,c-20000) .. c+20001, float64 sqrt distance, ring l holds
 * MAX[l] < dist <= MAX[l+1].  out[l*cap ...] receives ring l in scan order.
 * Returns -1-l if ring l is empty, -100 if cap is too small, else 0. */
int orc_ring_indexes(const int32_t *X, const int32_t *Y, uint32_t n, uint32_t centre, int levels,
                     uint32_t *out, uint32_t *counts, uint32_t cap) {
    const int64_t cx = X[centre], cy = Y[centre];
    uint32_t start = centre > MAX_SEARCH_AREA ? centre - MAX_SEARCH_AREA : 0;
    for (int l = 0; l < levels; ++l) counts[l] = 0;
    for (uint32_t i = start; i < n; ++i) {
        const int64_t dx = X[i] - cx, dy = Y[i] - cy;
        const double dist = sqrt((double)(dx * dx + dy * dy));
        for (int l = 0; l < levels; ++l) {
            if ((double)MAX_DISTS[l] < dist && dist <= (double)MAX_DISTS[l + 1]) {
                if (counts[l] >= cap) return -100;
                out[(size_t)l * cap + counts[l]++] = i;
            }
        }
        if ((uint64_t)i > (uint64_t)centre + MAX_SEARCH_AREA) break; /* tested after the append (:66-67) */
    }
    for (int l = 0; l < levels; ++l)
        if (counts[l] == 0) return -1 - l;
    return 0;
}

/* bcl_direct_reader.py:222-253 */
uint32_t orc_filter_offsets(const uint8_t *filt, uint32_t n, int32_t *off) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; ++i) off[i] = (filt[i] & 1) ? (int32_t)k++ : -1;
    return k;
}

/* one call -> code 0..3 = ACGT, 4 = N.
 * kind 1: BCL byte (bcl_direct_reader.py:347-361)
 * kind 2/3: CBCL nibble, 3 = non-PF wells excluded (:303-325) */
static inline uint8_t decode_call(const uint8_t *plane, int kind, uint32_t well, const int32_t *off) {
    if (kind == 1) {
        const uint8_t b = plane[well];
        return b ? (uint8_t)(b & 3) : 4;
    }
    int64_t w = well;
    if (kind == 3) {
        w = off[well];
        if (w == -1) return 4;
    }
    const uint8_t byte = plane[w / 2];
    const uint8_t nib = (w % 2) ? (uint8_t)(byte >> 4) : (uint8_t)(byte & 15);
    return nib ? (uint8_t)(nib & 3) : 4;
}

/* Tile.get_seqs on in-memory planes (:158-220): codes[i*len + p] */
void orc_get_codes(const uint8_t *const *planes, const int *kinds, int len, const uint8_t *filt, uint32_t n,
                   const int64_t *idx, uint32_t n_idx, uint8_t *codes, uint8_t *pf) {
    int32_t *off = NULL;
    for (int p = 0; p < len; ++p)
        if (kinds[p] == 3 && !off) {
            off = (int32_t *)malloc((size_t)n * sizeof(int32_t));
            orc_filter_offsets(filt, n, off);
        }
    for (uint32_t i = 0; i < n_idx; ++i) {
        for (int p = 0; p < len; ++p) codes[(size_t)i * len + p] = decode_call(planes[p], kinds[p], (uint32_t)idx[i], off);
        pf[i] = filt[idx[i]] & 1;
    }
    free(off);
}

/* unit-cost edit distance, common prefix/suffix stripped first */
int orc_levenshtein(const uint8_t *a, int la, const uint8_t *b, int lb) {
    while (la > 0 && lb > 0 && a[0] == b[0]) { ++a; ++b; --la; --lb; }
    while (la > 0 && lb > 0 && a[la - 1] == b[lb - 1]) { --la; --lb; }
    if (la == 0) return lb;
    if (lb == 0) return la;
    int stack_row[1100];
    int *row = lb + 1 <= 1100 ? stack_row : (int *)malloc((size_t)(lb + 1) * sizeof(int));
    for (int j = 0; j <= lb; ++j) row[j] = j;
    for (int i = 1; i <= la; ++i) {
        int diag = row[0];
        row[0] = i;
        for (int j = 1; j <= lb; ++j) {
            const int up = row[j];
            int v = diag + (a[i - 1] != b[j - 1]);
            if (up + 1 < v) v = up + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            diag = up;
            row[j] = v;
        }
    }
    const int r = row[lb];
    if (row != stack_row) free(row);
    return r;
}

/* Not a reference function: the definition of the exact prefix test the fused CUDA kernel stops reading a ring
 * well with (wd_seq.cuh, PrefixDP), restated on the full edit matrix so that tests can replay which plane bytes
 * the kernel needs.  After p symbols of b, against a of length la:
 *   min over |j - p| <= k, 0 <= j <= la of  D[j][p] + |j - p|,   D[j][p] = edit distance(a[0..j), b[0..p)).
 * A value > e proves Levenshtein(a, b) > e for equal-length strings and k = e / 2. */
int orc_prefix_band_min(const uint8_t *a, int la, const uint8_t *b, int p, int k) {
    int *col = (int *)malloc((size_t)(la + 1) * sizeof(int));
    for (int j = 0; j <= la; ++j) col[j] = j;                    /* D[j][0] */
    for (int c = 1; c <= p; ++c) {
        int diag = col[0];
        col[0] = c;
        for (int j = 1; j <= la; ++j) {
            const int left = col[j];
            int v = diag + (a[j - 1] != b[c - 1]);
            if (left + 1 < v) v = left + 1;
            if (col[j - 1] + 1 < v) v = col[j - 1] + 1;
            diag = left;
            col[j] = v;
        }
    }
    int best = 1 << 30;
    for (int j = (p - k > 0 ? p - k : 0); j <= (p + k < la ? p + k : la); ++j) {
        const int v = col[j] + (j > p ? j - p : p - j);
        if (v < best) best = v;
    }
    free(col);
    return best;
}

int orc_hamming(const uint8_t *a, const uint8_t *b, int n) {
    int d = 0;
    for (int i = 0; i < n; ++i) d += a[i] != b[i];
    return d;
}

/* main() inner loops + output_writer sums for one tile
 * (count_well_duplicates.py:228-265, :65-95).
 * per_target [t][1+2L] = valid, (dups, wells)*; counters [1+5L] = Targets,
 * (Wells, Dups, Hit, AccO, AccI)*; either may be NULL. */
void orc_count_tile(const uint8_t *const *planes, const int *kinds, int len, const uint8_t *filt, uint32_t n,
                    const uint32_t *centres, const uint32_t *level_offsets, const uint32_t *idx, uint32_t t,
                    int levels, int edit_distance, int use_hamming, int32_t *per_target, int64_t *counters) {
    int32_t *off = NULL;
    for (int p = 0; p < len; ++p)
        if (kinds[p] == 3 && !off) {
            off = (int32_t *)malloc((size_t)n * sizeof(int32_t));
            orc_filter_offsets(filt, n, off);
        }
    uint8_t *cseq = (uint8_t *)malloc((size_t)len + 1), *wseq = (uint8_t *)malloc((size_t)len + 1);
    const int width = 1 + 5 * levels, row = 1 + 2 * levels;
    if (counters) memset(counters, 0, (size_t)width * sizeof(int64_t));
    int *dups = (int *)malloc((size_t)levels * sizeof(int));
    for (uint32_t k = 0; k < t; ++k) {
        int32_t *pt = per_target ? per_target + (size_t)k * row : NULL;
        if (pt) memset(pt, 0, (size_t)row * sizeof(int32_t));
        const uint32_t c = centres[k];
        if (!(filt[c] & 1)) continue; /* :236-237 */
        for (int p = 0; p < len; ++p) cseq[p] = decode_call(planes[p], kinds[p], c, off);
        if (pt) pt[0] = 1;
        for (int l = 0; l < levels; ++l) {
            const uint32_t a = level_offsets[(size_t)k * levels + l], b = level_offsets[(size_t)k * levels + l + 1];
            int d = 0;
            for (uint32_t j = a; j < b; ++j) {
                for (int p = 0; p < len; ++p) wseq[p] = decode_call(planes[p], kinds[p], idx[j], off);
                const int dist = use_hamming ? orc_hamming(cseq, wseq, len) : orc_levenshtein(cseq, len, wseq, len);
                d += dist <= edit_distance;
            }
            dups[l] = d;
            if (pt) { pt[1 + 2 * l] = d; pt[2 + 2 * l] = (int32_t)(b - a); }
            if (counters) {
                counters[1 + 5 * l] += b - a;
                counters[2 + 5 * l] += d;
                counters[3 + 5 * l] += d > 0;
            }
        }
        if (counters) {
            counters[0] += 1;
            int seen = 0;
            for (int l = 0; l < levels; ++l) { seen |= dups[l] > 0; counters[4 + 5 * l] += seen; }
            seen = 0;
            for (int l = levels - 1; l >= 0; --l) { seen |= dups[l] > 0; counters[5 + 5 * l] += seen; }
        }
    }
    free(dups); free(cseq); free(wseq); free(off);
}
