// Packed base-call sequences and the distance predicates on them.
//
// A well's compared substring (<= 64*W symbols over {A,C,G,T,N}) is held as
// three bit-planes of W 64-bit words: lo = bit0 of the base, hi = bit1 of the
// base, nn = "is N".  Where nn is set, lo and hi are zero (canonical form), so
// symbol equality is plain bitwise equality of the three planes and "N is an
// ordinary character" (count_well_duplicates.py:238,251-252) comes for free.
// Bits at positions >= len are zero in all planes.
//
// The functions compile for the device (nvcc) and for the host (g++), the
// latter only so tests/ can drive them exhaustively against the oracle without
// a GPU (tests/cpu_seq_harness.cpp); the library never calls them on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WD_HD __host__ __device__ __forceinline__
#else
#define WD_HD inline
#endif

namespace wd {

WD_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

template <int W>
struct PSeq {
    uint64_t lo[W], hi[W], nn[W];
};

// valid-bit mask of word w for a sequence of len symbols
WD_HD uint64_t len_mask(int len, int w) {
    int rem = len - 64 * w;
    if (rem >= 64) return ~0ull;
    if (rem <= 0) return 0ull;
    return (1ull << rem) - 1ull;
}

template <int W>
WD_HD void pseq_clear(PSeq<W> &s) {
#pragma unroll
    for (int w = 0; w < W; ++w) s.lo[w] = s.hi[w] = s.nn[w] = 0ull;
}

// code: 0..3 = A,C,G,T; 4 = N
template <int W>
WD_HD void pseq_set(PSeq<W> &s, int pos, unsigned code) {
    const int w = pos >> 6;
    const int b = pos & 63;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == w) {
            s.lo[i] |= (uint64_t)(code & 1u) << b;
            s.hi[i] |= (uint64_t)((code >> 1) & 1u) << b;
            s.nn[i] |= (uint64_t)(code >> 2) << b;
        }
    }
}

template <int W>
WD_HD unsigned pseq_get(const PSeq<W> &s, int pos) {
    const int w = pos >> 6;
    const int b = pos & 63;
    uint64_t lo = 0, hi = 0, nn = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == w) { lo = s.lo[i]; hi = s.hi[i]; nn = s.nn[i]; }
    }
    return (unsigned)((lo >> b) & 1u) | ((unsigned)((hi >> b) & 1u) << 1) | ((unsigned)((nn >> b) & 1u) << 2);
}

// Positional mismatches (Levenshtein.hamming on equal-length strings,
// count_well_duplicates.py:200).
template <int W>
WD_HD int hamming(const PSeq<W> &a, const PSeq<W> &b) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w)
        d += popc64((a.lo[w] ^ b.lo[w]) | (a.hi[w] ^ b.hi[w]) | (a.nn[w] ^ b.nn[w]));
    return d;
}

// word w of the W-word bit string p shifted so that result[j] = p[j + d]
// (zeros shifted in); |d| < 64.
template <int W>
WD_HD uint64_t shifted_word(const uint64_t *p, int w, int d) {
    if (d == 0) return p[w];
    if (d > 0) {
        uint64_t v = p[w] >> d;
        if (w + 1 < W) v |= p[w + 1] << (64 - d);
        return v;
    }
    const int s = -d;
    uint64_t v = p[w] << s;
    if (w > 0) v |= p[w - 1] >> (64 - s);
    return v;
}

// Shifted-Hamming lower bound.  In any edit script of cost <= e between two
// strings of equal length, #insertions == #deletions <= k = e/2, so a symbol
// of b that the script matches exactly is matched to a[j+d] for some
// |d| <= k; every other symbol of b costs at least one edit.  Hence the
// number of positions j with a[j+d] != b[j] for ALL |d| <= k is <= e whenever
// Lev(a,b) <= e -- and that stays true if only the positions j < prefix are
// counted, which is what lets the gather stop reading planes for a well as
// soon as its first symbols rule it out.  a is known over [0, len); b over
// [0, prefix).  Returns the count.
template <int W>
WD_HD int shd_bad_count(const PSeq<W> &a, const PSeq<W> &b, int len, int prefix, int k) {
    int bad = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint64_t lm = len_mask(prefix, w);
        if (lm != 0ull) {
            uint64_t all = lm;
            for (int d = -k; d <= k; ++d) {
                uint64_t mm = (shifted_word<W>(a.lo, w, d) ^ b.lo[w]) |
                              (shifted_word<W>(a.hi, w, d) ^ b.hi[w]) |
                              (shifted_word<W>(a.nn, w, d) ^ b.nn[w]);
                // positions whose partner j+d falls outside [0, len) cannot match
                if (d > 0) {
                    const int cut = len - d - 64 * w;   // first invalid bit in this word
                    if (cut <= 0) mm = ~0ull;
                    else if (cut < 64) mm |= ~((1ull << cut) - 1ull);
                } else if (d < 0) {
                    const int cut = -d - 64 * w;        // bits [0, cut) invalid
                    if (cut >= 64) mm = ~0ull;
                    else if (cut > 0) mm |= (1ull << cut) - 1ull;
                }
                all &= mm;
            }
            bad += popc64(all);
        }
    }
    return bad;
}

// true when the bound proves Lev(a,b) > e
template <int W>
WD_HD bool shd_rejects(const PSeq<W> &a, const PSeq<W> &b, int len, int e) {
    const int k = e >> 1;
    if (k >= 64) return false;
    return shd_bad_count<W>(a, b, len, len, k) > e;
}

// true when the first `prefix` symbols of b (the rest of b still unknown,
// stored as zero bits) already prove dist(a, b) > e under the chosen metric.
template <int W>
WD_HD bool prefix_rejects(const PSeq<W> &a, const PSeq<W> &b, int len, int prefix, int e, bool use_hamming) {
    if (e < 0) return true;
    if (e >= len) return false;
    if (use_hamming || e < 2) {                // Lev <= 1 <=> Ham <= 1
        int d = 0;
#pragma unroll
        for (int w = 0; w < W; ++w)
            d += popc64(((a.lo[w] ^ b.lo[w]) | (a.hi[w] ^ b.hi[w]) | (a.nn[w] ^ b.nn[w])) & len_mask(prefix, w));
        return d > e;
    }
    const int k = e >> 1;
    if (k >= 64) return false;
    return shd_bad_count<W>(a, b, len, prefix, k) > e;
}

// Exact unit-cost edit distance between two equal-length packed strings
// (Levenshtein.distance, count_well_duplicates.py:200,252): Myers' bit-vector
// algorithm in Hyyro's global-alignment formulation, block-wise over W words
// with the horizontal delta carried between blocks.
template <int W>
WD_HD int myers_distance(const PSeq<W> &a, const PSeq<W> &b, int len) {
    if (len <= 0) return 0;
    uint64_t Pv[W], Mv[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
    const int last_w = (len - 1) >> 6;
    const uint64_t last_bit = 1ull << ((len - 1) & 63);
    int score = len;
    for (int j = 0; j < len; ++j) {
        const unsigned c = pseq_get<W>(b, j);
        const uint64_t tlo = (c & 1u) ? ~0ull : 0ull;
        const uint64_t thi = (c & 2u) ? ~0ull : 0ull;
        const uint64_t tn = (c & 4u) ? ~0ull : 0ull;
        int hin = 1;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (w <= last_w) {
                uint64_t Eq = (tn & a.nn[w]) | (~tn & ~a.nn[w] & ~(a.lo[w] ^ tlo) & ~(a.hi[w] ^ thi));
                const uint64_t pv = Pv[w], mv = Mv[w];
                const uint64_t Xv = Eq | mv;
                if (hin < 0) Eq |= 1ull;
                const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                uint64_t Ph = mv | ~(Xh | pv);
                uint64_t Mh = pv & Xh;
                const uint64_t top = (w == last_w) ? last_bit : (1ull << 63);
                int hout = 0;
                if (Ph & top) hout = 1;
                else if (Mh & top) hout = -1;
                Ph <<= 1;
                Mh <<= 1;
                if (hin < 0) Mh |= 1ull;
                else if (hin > 0) Ph |= 1ull;
                Pv[w] = Mh | ~(Xv | Ph);
                Mv[w] = Ph & Xv;
                hin = hout;
            }
        }
        score += hin;
    }
    return score;
}

// Incremental form of the same dynamic programme, for the fused gather: b is
// fed one symbol at a time (one plane byte at a time), and after p symbols the
// column D[.][p] of the edit matrix against a is available as vertical deltas,
//   D[j][p] = p + popc(Pv & below(j)) - popc(Mv & below(j)).
// Lev(a, b) <= e on equal-length strings needs an alignment that crosses
// column p at some row j with |j - p| <= k = e/2 (it has to spend |j - p|
// indels before the crossing and as many after it), and costs at least
// D[j][p] + |j - p|.  So  min over |j-p| <= k of D[j][p] + |j - p|  > e
// proves dist > e from the first p symbols of b and the first p + k of a --
// the exact prefix test; at p = len it is the distance test itself.  Rows
// below the band never carry an alignment of cost <= e, so a needs to be known
// only up to row p + k (unknown rows count as mismatches), and words of the
// bit-vectors wholly below the band are left in their initial state until the
// band reaches them (their cells then over-estimate D, which cannot create a
// value <= e).
// The vectors are kept as 32-bit words (two per 64-symbol word of a): 32-bit
// adds and logic ops are single instructions on the GPU, and with the lazy
// activation above the first rounds of a 50-symbol comparison touch one word.
WD_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// valid-bit mask of 32-bit word w for a prefix of n symbols
WD_HD uint32_t len_mask32(int n, int w) {
    const int rem = n - 32 * w;
    if (rem >= 32) return ~0u;
    if (rem <= 0) return 0u;
    return (1u << rem) - 1u;
}

template <int W>
struct PrefixDP {
    uint32_t Pv[2 * W], Mv[2 * W];
};

template <int W>
WD_HD void pdp_init(PrefixDP<W> &s) {
#pragma unroll
    for (int w = 0; w < 2 * W; ++w) { s.Pv[w] = ~0u; s.Mv[w] = 0u; }
}

// consume b[p] = c (0..3 bases, 4 = N, anything else matches nothing); a is
// valid over [0, known_a), known_a >= min(len, p + k + 1)
template <int W>
WD_HD void pdp_step(PrefixDP<W> &s, const PSeq<W> &a, int known_a, int len, int p, int k, unsigned c) {
    const uint32_t tlo = (c & 1u) ? ~0u : 0u;
    const uint32_t thi = (c & 2u) ? ~0u : 0u;
    const uint32_t tn = (c & 4u) ? ~0u : 0u;
    const uint32_t tany = c <= 4u ? ~0u : 0u;
    const int last_w = (len - 1) >> 5;
    const int band_w = (p + k + 1) >> 5;
    int hin = 1;                                  // D[0][p+1] - D[0][p]
#pragma unroll
    for (int w = 0; w < 2 * W; ++w) {
        if (w <= last_w && w <= band_w) {
            const uint32_t alo = (uint32_t)(a.lo[w >> 1] >> (32 * (w & 1)));
            const uint32_t ahi = (uint32_t)(a.hi[w >> 1] >> (32 * (w & 1)));
            const uint32_t ann = (uint32_t)(a.nn[w >> 1] >> (32 * (w & 1)));
            // canonical form (N carries no base bits): equal symbols <=> equal in all three planes
            uint32_t Eq = ~((alo ^ tlo) | (ahi ^ thi) | (ann ^ tn));
            Eq &= len_mask32(known_a, w) & tany;
            const uint32_t pv = s.Pv[w], mv = s.Mv[w];
            const uint32_t Xv = Eq | mv;
            if (hin < 0) Eq |= 1u;
            const uint32_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            uint32_t Ph = mv | ~(Xh | pv);
            uint32_t Mh = pv & Xh;
            int hout = 0;
            if (Ph >> 31) hout = 1;
            else if (Mh >> 31) hout = -1;
            Ph <<= 1;
            Mh <<= 1;
            if (hin < 0) Mh |= 1u;
            else if (hin > 0) Ph |= 1u;
            s.Pv[w] = Mh | ~(Xv | Ph);
            s.Mv[w] = Ph & Xv;
            hin = hout;
        }
    }
}

// after p symbols of b:  min over |j - p| <= k, 0 <= j <= len  of  D[j][p] + |j - p|
template <int W>
WD_HD int pdp_band_min(const PrefixDP<W> &s, int len, int p, int k) {
    const int jlo = p - k > 0 ? p - k : 0;
    const int jhi = p + k < len ? p + k : len;
    int d = p;
#pragma unroll
    for (int w = 0; w < 2 * W; ++w) {
        const uint32_t m = len_mask32(jlo, w);
        d += popc32(s.Pv[w] & m) - popc32(s.Mv[w] & m);
    }
    int best = d + (p - jlo);
    for (int j = jlo; j < jhi; ++j) {             // row j -> j + 1 is bit j
        const int w = j >> 5, b = j & 31;
        uint32_t pv = 0, mv = 0;
#pragma unroll
        for (int i = 0; i < 2 * W; ++i)
            if (i == w) { pv = s.Pv[i]; mv = s.Mv[i]; }
        d += (int)((pv >> b) & 1u) - (int)((mv >> b) & 1u);
        const int off = j + 1 - p;
        const int v = d + (off < 0 ? -off : off);
        best = v < best ? v : best;
    }
    return best;
}

// The same two operations for the rounds whose band lies inside the first 32
// rows (p + n + k + 1 <= 32 for a round of n symbols starting at p): only word
// 0 of the vectors is active, row 0 always enters with +1, and the word-select
// loops and their branches of the general form fall away -- a third of the
// instructions.  This is the first round of every ring well, i.e. nearly all
// the symbols the fused kernel ever feeds.  Bit-identical to the general form
// (tests/test_seq_predicates_cpu.py drives both through the same dispatch).
WD_HD bool pdp_round_in_word0(int p, int n, int k) { return p + n + k + 1 <= 32; }

// alo/ahi/ann: word 0 of a's planes; amask: rows of word 0 that are known (len_mask32(known_a, 0))
template <int W>
WD_HD void pdp_step_word0(PrefixDP<W> &s, uint32_t alo, uint32_t ahi, uint32_t ann, uint32_t amask, unsigned c) {
    const uint32_t tlo = (c & 1u) ? ~0u : 0u;
    const uint32_t thi = (c & 2u) ? ~0u : 0u;
    const uint32_t tn = (c & 4u) ? ~0u : 0u;
    const uint32_t tany = c <= 4u ? ~0u : 0u;
    const uint32_t Eq = ~((alo ^ tlo) | (ahi ^ thi) | (ann ^ tn)) & amask & tany;
    const uint32_t pv = s.Pv[0], mv = s.Mv[0];
    const uint32_t Xv = Eq | mv;
    const uint32_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
    const uint32_t Ph = ((mv | ~(Xh | pv)) << 1) | 1u;          // row 0: D[0][p+1] - D[0][p] = +1
    const uint32_t Mh = (pv & Xh) << 1;
    s.Pv[0] = Mh | ~(Xv | Ph);
    s.Mv[0] = Ph & Xv;
}

// the same step with the Eq mask of the symbol handed in (the fused kernel keeps one per symbol and centre in
// shared memory: Eq depends on the centre and the symbol only, not on the ring well)
template <int W>
WD_HD void pdp_step_word0_eq(PrefixDP<W> &s, uint32_t Eq) {
    const uint32_t pv = s.Pv[0], mv = s.Mv[0];
    const uint32_t Xv = Eq | mv;
    const uint32_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
    const uint32_t Ph = (mv | ~(Xh | pv)) * 2u + 1u;            // << 1, and row 0: D[0][p+1] - D[0][p] = +1 (one multiply-add)
    const uint32_t Mh = (pv & Xh) << 1;
    s.Pv[0] = Mh | ~(Xv | Ph);
    s.Mv[0] = Ph & Xv;
}

// pdp_band_min for p + k <= 32 (all rows of the band in word 0)
template <int W>
WD_HD int pdp_band_min_word0(const PrefixDP<W> &s, int len, int p, int k) {
    const uint32_t pv = s.Pv[0], mv = s.Mv[0];
    if (k == 1 && p >= 1 && p < len) {
        // the common band (e = 2 or 3): rows p-1, p, p+1 without a loop
        const uint32_t below = (1u << (p - 1)) - 1u;
        const int d0 = p + popc32(pv & below) - popc32(mv & below);                       // D[p-1][p]
        const int d1 = d0 + (int)((pv >> (p - 1)) & 1u) - (int)((mv >> (p - 1)) & 1u);     // D[p][p]
        const int d2 = d1 + (int)((pv >> p) & 1u) - (int)((mv >> p) & 1u);                 // D[p+1][p]
        const int side = (d0 < d2 ? d0 : d2) + 1;
        return d1 < side ? d1 : side;
    }
    const int jlo = p - k > 0 ? p - k : 0;
    const int jhi = p + k < len ? p + k : len;
    const uint32_t below = jlo >= 32 ? ~0u : ((1u << jlo) - 1u);
    int d = p + popc32(pv & below) - popc32(mv & below);
    int best = d + (p - jlo);
    for (int j = jlo; j < jhi; ++j) {             // row j -> j + 1 is bit j
        d += (int)((pv >> j) & 1u) - (int)((mv >> j) & 1u);
        const int off = j + 1 - p;
        const int v = d + (off < 0 ? -off : off);
        best = v < best ? v : best;
    }
    return best;
}

// Cheap necessary condition for dist(a, b) <= e from the first 32 symbols of b
// (exhaustive mode, wd_exhaustive.cu), on the lo/hi planes only: N aliases to A,
// so every true match stays a match and the count of unmatched positions can
// only shrink (see shd_bad_count).  In an edit script of cost <= e on
// equal-length strings every exactly matched symbol b[j] is matched to
// a[j + d] with |d| <= k = e/2 (k = 0 for Hamming and for e < 2); all other
// positions of b cost an edit each.  Per string a, the set
// S_j = {a[j + d] : |d| <= k} is kept as four 32-bit planes (one per base), so
//   unmatched(b) = ~mux(b.hi, mux(b.lo, SA, SC), mux(b.lo, SG, ST))
// costs three logic instructions whatever k is; popc(unmatched) <= e is the
// test.  Partners outside [0, len) read as A (zero bits): that only weakens it.
struct Head32Sets {
    uint32_t sa, sc, sg, st;
    WD_HD void clear() { sa = sc = sg = st = 0u; }
    // lo, hi: first 64-symbol word of a's base planes; k <= 31
    WD_HD void set(uint64_t lo, uint64_t hi, int k) {
        clear();
        for (int d = -k; d <= k; ++d) {
            const uint32_t l = d >= 0 ? (uint32_t)(lo >> d) : (uint32_t)lo << (-d);
            const uint32_t h = d >= 0 ? (uint32_t)(hi >> d) : (uint32_t)hi << (-d);
            sa |= ~l & ~h; sc |= l & ~h; sg |= ~l & h; st |= l & h;
        }
    }
    // blo, bhi: base planes of b's first 32 symbols
    // positions of b that find a partner (three 3-input logic ops; the complement would cost a fourth)
    WD_HD uint32_t matched(uint32_t blo, uint32_t bhi) const {
        const uint32_t t0 = (blo & sc) | (~blo & sa);
        const uint32_t t1 = (blo & st) | (~blo & sg);
        return (bhi & t1) | (~bhi & t0);
    }
    WD_HD uint32_t unmatched(uint32_t blo, uint32_t bhi) const { return ~matched(blo, bhi); }
};

// k of the test above for (e, metric): Levenshtein <= 1 <=> Hamming <= 1 on equal lengths
WD_HD int head32_k(int e, bool use_hamming) {
    if (use_hamming || e < 2) return 0;
    return (e >> 1) < 31 ? (e >> 1) : 31;
}

// dist(a, b) <= e under the reference's chosen metric.
template <int W>
WD_HD bool is_duplicate(const PSeq<W> &a, const PSeq<W> &b, int len, int e, bool use_hamming) {
    if (e < 0) return false;
    const int ham = hamming<W>(a, b);
    if (ham <= e) return true;                 // Lev <= Ham
    if (use_hamming) return false;
    if (e < 2) return false;                   // an indel pair costs 2: Lev <= 1 <=> Ham <= 1
    if (e >= len) return true;                 // len substitutions always suffice
    if (shd_rejects<W>(a, b, len, e)) return false;
    return myers_distance<W>(a, b, len) <= e;
}

// the value the reference logs for a duplicate pair (count_well_duplicates.py:262)
template <int W>
WD_HD int exact_distance(const PSeq<W> &a, const PSeq<W> &b, int len, bool use_hamming) {
    const int ham = hamming<W>(a, b);
    if (use_hamming || ham <= 2) return ham;   // beating Ham needs an indel pair (2) replacing >= 3 mismatches
    return myers_distance<W>(a, b, len);
}

}  // namespace wd
