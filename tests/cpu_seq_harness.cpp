// TEST-ONLY host build of well_duplicates_b200/csrc/wd_seq.cuh so the packed
// distance predicates can be checked against the oracle without a GPU.
// Never linked into libwelldup.so.
#include <stdint.h>
#include "../well_duplicates_b200/csrc/wd_seq.cuh"

template <int W>
static void run(const uint8_t *a, const uint8_t *b, int len, int e, int ham, int *dup, int *exact, int *shd) {
    wd::PSeq<W> pa, pb;
    wd::pseq_clear(pa);
    wd::pseq_clear(pb);
    for (int i = 0; i < len; ++i) {
        wd::pseq_set<W>(pa, i, a[i]);
        wd::pseq_set<W>(pb, i, b[i]);
    }
    *dup = wd::is_duplicate<W>(pa, pb, len, e, ham != 0) ? 1 : 0;
    *exact = ham ? wd::hamming<W>(pa, pb) : wd::myers_distance<W>(pa, pb, len);
    *shd = wd::shd_rejects<W>(pa, pb, len, e) ? 1 : 0;
}

// smallest multiple-of-8 prefix of b at which prefix_rejects() fires (0 = never)
template <int W>
static int first_reject(const uint8_t *a, const uint8_t *b, int len, int e, int ham) {
    wd::PSeq<W> pa, pb;
    wd::pseq_clear(pa);
    wd::pseq_clear(pb);
    for (int i = 0; i < len; ++i) wd::pseq_set<W>(pa, i, a[i]);
    for (int k = 0; k < len;) {
        const int upto = k + 8 < len ? k + 8 : len;
        for (; k < upto; ++k) wd::pseq_set<W>(pb, k, b[k]);
        if (wd::prefix_rejects<W>(pa, pb, len, k, e, ham != 0)) return k;
    }
    return 0;
}

extern "C" int seq_first_reject(const uint8_t *a, const uint8_t *b, int len, int words, int e, int ham) {
    switch (words) {
        case 1: return first_reject<1>(a, b, len, e, ham);
        case 2: return first_reject<2>(a, b, len, e, ham);
        case 4: return first_reject<4>(a, b, len, e, ham);
        case 8: return first_reject<8>(a, b, len, e, ham);
        case 16: return first_reject<16>(a, b, len, e, ham);
    }
    return -1;
}

extern "C" int seq_check(const uint8_t *a, const uint8_t *b, int len, int words, int e, int ham,
                         int *dup, int *exact, int *shd) {
    switch (words) {
        case 1: run<1>(a, b, len, e, ham, dup, exact, shd); return 0;
        case 2: run<2>(a, b, len, e, ham, dup, exact, shd); return 0;
        case 4: run<4>(a, b, len, e, ham, dup, exact, shd); return 0;
        case 8: run<8>(a, b, len, e, ham, dup, exact, shd); return 0;
        case 16: run<16>(a, b, len, e, ham, dup, exact, shd); return 0;
    }
    return -1;
}

// Incremental prefix DP (PrefixDP): feeds b symbol by symbol with a known only
// up to p + k, as the fused kernel does.  out[p] (p = 1..len) = pdp_band_min
// after p symbols.
template <int W>
static void prefix_dp(const uint8_t *a, const uint8_t *b, int len, int k, int chunk, int *out) {
    wd::PSeq<W> pa;
    wd::pseq_clear(pa);
    int known = 0;
    wd::PrefixDP<W> s;
    wd::pdp_init(s);
    for (int p = 0; p < len; ++p) {
        int need = p + k + 1 < len ? p + k + 1 : len;
        while (known < need) {                    // a arrives in chunks, like the centre in the kernel
            const int upto = known + chunk < len ? known + chunk : len;
            for (; known < upto; ++known) wd::pseq_set<W>(pa, known, a[known]);
        }
        // the kernel's dispatch: rounds inside the first 32 rows take the word-0 forms
        if (wd::pdp_round_in_word0(p, 1, k)) {
            wd::PrefixDP<W> s2 = s;
            wd::pdp_step_word0<W>(s, (uint32_t)pa.lo[0], (uint32_t)pa.hi[0], (uint32_t)pa.nn[0], wd::len_mask32(known, 0), b[p]);
            // ... and with the Eq mask looked up per symbol, as the kernel keeps it in shared memory
            uint32_t eq[5];
            for (unsigned c = 0; c < 5; ++c) {
                const uint32_t tlo = (c & 1u) ? ~0u : 0u, thi = (c & 2u) ? ~0u : 0u, tn = (c & 4u) ? ~0u : 0u;
                eq[c] = ~(((uint32_t)pa.lo[0] ^ tlo) | ((uint32_t)pa.hi[0] ^ thi) | ((uint32_t)pa.nn[0] ^ tn)) & wd::len_mask32(known, 0);
            }
            wd::pdp_step_word0_eq<W>(s2, eq[b[p]]);
            if (s2.Pv[0] != s.Pv[0] || s2.Mv[0] != s.Mv[0]) { out[p + 1] = -2000; continue; }
            out[p + 1] = wd::pdp_band_min_word0<W>(s, len, p + 1, k);
            if (out[p + 1] != wd::pdp_band_min<W>(s, len, p + 1, k)) out[p + 1] = -1000;     // the two forms must agree
        } else {
            wd::pdp_step<W>(s, pa, known, len, p, k, b[p]);
            out[p + 1] = wd::pdp_band_min<W>(s, len, p + 1, k);
        }
    }
}

extern "C" int seq_prefix_dp(const uint8_t *a, const uint8_t *b, int len, int words, int k, int chunk, int *out) {
    switch (words) {
        case 1: prefix_dp<1>(a, b, len, k, chunk, out); return 0;
        case 2: prefix_dp<2>(a, b, len, k, chunk, out); return 0;
        case 4: prefix_dp<4>(a, b, len, k, chunk, out); return 0;
        case 8: prefix_dp<8>(a, b, len, k, chunk, out); return 0;
        case 16: prefix_dp<16>(a, b, len, k, chunk, out); return 0;
    }
    return -1;
}

template <int W>
static int head32(const uint8_t *a, const uint8_t *b, int len, int e, int ham) {
    wd::PSeq<W> pa, pb;
    wd::pseq_clear(pa);
    wd::pseq_clear(pb);
    for (int i = 0; i < len; ++i) {
        wd::pseq_set<W>(pa, i, a[i]);
        wd::pseq_set<W>(pb, i, b[i]);
    }
    if (e < 0) return 1;
    if (e >= len) return 0;                      // the kernel does not compare sequences then
    wd::Head32Sets s;
    s.set(pa.lo[0], pa.hi[0], wd::head32_k(e, ham != 0));
    return wd::popc32(s.unmatched((uint32_t)pb.lo[0], (uint32_t)pb.hi[0])) > e ? 1 : 0;
}

// 1 when the 32-symbol pre-filter of exhaustive mode rejects the pair
extern "C" int seq_head32_rejects(const uint8_t *a, const uint8_t *b, int len, int words, int e, int ham) {
    switch (words) {
        case 1: return head32<1>(a, b, len, e, ham);
        case 2: return head32<2>(a, b, len, e, ham);
        case 4: return head32<4>(a, b, len, e, ham);
        case 8: return head32<8>(a, b, len, e, ham);
        case 16: return head32<16>(a, b, len, e, ham);
    }
    return -1;
}
