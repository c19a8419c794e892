// Host side of the staging pipeline: gunzip of .bcl.gz planes and CBCL tile
// blocks straight into the (pinned) plane buffers the kernels read.
//
// BASELINE.json's north_star keeps gunzip on the host; SURVEY 8(f2) names it as
// the wall clock of a run from compressed files (gzip.open(...).read() per tile
// and cycle in bcl_direct_reader.py:207-208, :300-301, ~0.1 GB/s per core with
// zlib on base-call bytes).  This file replaces that step with
//   * an inflate written for this data: 64-bit bit buffer refilled branch-free,
//     a 10-bit first-level literal/length table whose entries deliver two literals,
//     or a whole short match (length + distance code), per lookup where the codes
//     fit the index; the next entry is looked up before the match copy; matches
//     are copied in 16-byte pieces;
//   * CRC-32 by carry-less multiplication (PCLMULQDQ folding), table fallback;
//   * a job list executed by a pool of native threads (file read + inflate +
//     check, no Python in the loop), writing each member where the caller says.
// RFC 1951 / RFC 1952 semantics, with gzip.open()'s behaviour for multi-member
// files, trailing zero padding, truncated input and CRC / length mismatches.
#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/welldup.h"

namespace wd {
void set_error(const char *fmt, ...);   // wd_api.cu: thread-local message behind wd_last_error()
}

namespace {

// ---------------------------------------------------------------------------------------------
// CRC-32 (IEEE 802.3, reflected, as RFC 1952 section 8)
// ---------------------------------------------------------------------------------------------
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};
const CrcTables g_crc;

// state is the raw register (already inverted by the caller)
uint32_t crc32_tables(uint32_t c, const uint8_t *p, size_t n) {
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = g_crc.t[7][w & 0xff] ^ g_crc.t[6][(w >> 8) & 0xff] ^ g_crc.t[5][(w >> 16) & 0xff] ^
            g_crc.t[4][(w >> 24) & 0xff] ^ g_crc.t[3][(w >> 32) & 0xff] ^ g_crc.t[2][(w >> 40) & 0xff] ^
            g_crc.t[1][(w >> 48) & 0xff] ^ g_crc.t[0][w >> 56];
        p += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ g_crc.t[0][(c ^ *p++) & 0xff];
    return c;
}

#if defined(__x86_64__)
// Folding by carry-less multiplication (PCLMULQDQ).  The buffer is read as 128-bit blocks B0 B1 ...; an
// accumulator A stands for a prefix of the message as a polynomial mod P, and moving it D bits further
// along is a multiplication by x^D mod P: with A = (lo, hi) the two halves are multiplied separately,
//     A * x^D  ==  lo * (x^(D+32) mod P)  ^  hi * (x^(D-32) mod P)        (mod P)
// (the +-32 comes from the 64-bit halves sitting 64 bits apart around the 32-bit remainder, and the
// product of two bit-reflected 64-bit values comes out one bit low, hence the << 1 below).  Four
// accumulators walk the buffer 64 bytes at a time (D = 512), are merged with D = 128, the leftover
// 16-byte blocks are folded in with D = 128 as well -- and what remains is a 16-byte message with the
// same CRC register as the whole buffer, which the table loop finishes.  The multipliers are computed
// from the polynomial when first needed, not tabulated.
struct FoldKeys {
    __m128i by512, by128;
};

// x^n mod P in the register's (bit-reflected) order: bit 31 is x^0
static uint32_t xpow_mod_p(unsigned n) {
    uint32_t r = 0x80000000u;
    for (unsigned i = 0; i < n; ++i) r = (r >> 1) ^ ((r & 1u) ? 0xEDB88320u : 0u);
    return r;
}

static const FoldKeys &fold_keys() {
    static const FoldKeys k = [] {
        auto mult = [](unsigned n) { return (long long)((uint64_t)xpow_mod_p(n) << 1); };
        FoldKeys f;
        f.by512 = _mm_set_epi64x(mult(512 - 32), mult(512 + 32));       // high half, low half
        f.by128 = _mm_set_epi64x(mult(128 - 32), mult(128 + 32));
        return f;
    }();
    return k;
}

__attribute__((target("pclmul,sse4.1")))
static inline __m128i fold_into(__m128i acc, __m128i key, __m128i next) {
    const __m128i lo = _mm_clmulepi64_si128(acc, key, 0x00);
    const __m128i hi = _mm_clmulepi64_si128(acc, key, 0x11);
    return _mm_xor_si128(_mm_xor_si128(lo, hi), next);
}

__attribute__((target("pclmul,sse4.1")))
uint32_t crc32_clmul(uint32_t c, const uint8_t *p, size_t n) {   // n >= 64, n % 16 == 0
    const FoldKeys &k = fold_keys();
    const __m128i *blk = reinterpret_cast<const __m128i *>(p);
    size_t left = n / 16;
    __m128i acc[4];
    for (int j = 0; j < 4; ++j) acc[j] = _mm_loadu_si128(blk + j);
    acc[0] = _mm_xor_si128(acc[0], _mm_cvtsi32_si128((int)c));      // the running register enters with the first 4 bytes
    blk += 4;
    left -= 4;
    for (; left >= 4; blk += 4, left -= 4)
        for (int j = 0; j < 4; ++j) acc[j] = fold_into(acc[j], k.by512, _mm_loadu_si128(blk + j));
    __m128i a = acc[0];
    for (int j = 1; j < 4; ++j) a = fold_into(a, k.by128, acc[j]);
    for (; left > 0; ++blk, --left) a = fold_into(a, k.by128, _mm_loadu_si128(blk));
    uint8_t tail[16];
    _mm_storeu_si128(reinterpret_cast<__m128i *>(tail), a);
    return crc32_tables(0u, tail, sizeof(tail));
}

bool have_clmul() {
    static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}
#endif

uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n) {
    uint32_t c = ~crc;
#if defined(__x86_64__)
    if (n >= 128 && have_clmul()) {
        const size_t body = n & ~(size_t)15;
        c = crc32_clmul(c, p, body);
        p += body;
        n -= body;
    }
#endif
    return ~crc32_tables(c, p, n);
}

// ---------------------------------------------------------------------------------------------
// DEFLATE (RFC 1951)
// ---------------------------------------------------------------------------------------------
// Decode-table entry (32 bits):
//   bits 0..4    bits to drop from the bit buffer (code length [+ extra bits]; both codes of a
//                literal pair; length code + distance code + distance extra bits of a fused match;
//                for a pointer: the first-level width)
//   bits 5..7    kind
//   bits 8..11   where the extra bits start = code length (base + extra entries; both codes of a
//                fused match) / code length of the first literal / second-level width (pointer)
//   bits 12..15  bytes the entry produces: 1 literal, 2 literals (the second one in bits 24..31: two
//                short codes decoded by one lookup), length 3..10 of a fused match
//   bits 16..31  literal byte(s) / base value / index of the second-level table
enum : uint32_t {
    K_LITERAL = 0u << 5, K_BASE = 1u << 5, K_POINTER = 2u << 5, K_SPECIAL = 3u << 5,
    K_MATCH = 4u << 5,          // length code without extra bits + whole distance code in one first-level entry
    K_MASK = 7u << 5
};
#define NB(e) ((e) & 31u)
constexpr uint32_t ENTRY_INVALID = K_SPECIAL | (1u << 16) | 1u;   // drops one bit, never reached twice
constexpr uint32_t ENTRY_EOB_PAYLOAD = 0;

constexpr int LITLEN_BITS = 10;      // 10 vs 11 vs 12 measured on base-call planes: table build per block outweighs the extra pairs
constexpr int DIST_BITS = 8;
constexpr int PRE_BITS = 7;
constexpr int LITLEN_SYMS = 288;
constexpr int DIST_SYMS = 32;
// first level + second-level tables: every long code owns at most one second-level table of
// 2^(15 - first level) entries (a loose bound; build_table checks it anyway)
constexpr int LITLEN_TABLE = (1 << LITLEN_BITS) + 288 * (1 << (15 - LITLEN_BITS));
constexpr int DIST_TABLE = (1 << DIST_BITS) + 32 * (1 << (15 - DIST_BITS));

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum { INF_OK = 0, INF_TRUNCATED = 1, INF_BAD_DATA = 2, INF_OUT_FULL = 3 };

enum TableKind { T_PRECODE, T_LITLEN, T_DIST };

inline uint32_t symbol_entry(TableKind kind, int sym, int len) {
    switch (kind) {
    case T_PRECODE:
        return K_LITERAL | ((uint32_t)sym << 16) | (uint32_t)len;
    case T_LITLEN:
        if (sym < 256) return K_LITERAL | ((uint32_t)sym << 16) | (1u << 12) | ((uint32_t)len << 8) | (uint32_t)len;
        if (sym == 256) return K_SPECIAL | (ENTRY_EOB_PAYLOAD << 16) | (uint32_t)len;
        if (sym > 285) return K_SPECIAL | (1u << 16) | (uint32_t)len;       // 286, 287: never valid in data
        return K_BASE | ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)len << 8) | (uint32_t)(len + LEN_EXTRA[sym - 257]);
    default:
        if (sym > 29) return K_SPECIAL | (1u << 16) | (uint32_t)len;        // 30, 31: never valid in data
        return K_BASE | ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)len << 8) | (uint32_t)(len + DIST_EXTRA[sym]);
    }
}

inline uint32_t bit_reverse(uint32_t code, int len) {      // len <= 15
    uint32_t v = code;
    v = ((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u);
    v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u);
    v = ((v & 0x0f0fu) << 4) | ((v >> 4) & 0x0f0fu);
    v = ((v & 0x00ffu) << 8) | ((v >> 8) & 0x00ffu);
    return v >> (16 - len);
}

// Canonical Huffman code -> two-level lookup table indexed by the next bits of
// the stream (LSB first).  Returns false for an over-subscribed code, or an
// incomplete one other than the single-code case zlib accepts.
bool build_table(TableKind kind, const uint8_t *lens, int n_syms, int first_bits, uint32_t *table, int table_cap) {
    int count[16] = {0};
    for (int s = 0; s < n_syms; ++s) count[lens[s]]++;
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    const int first_size = 1 << first_bits;
    if (max_len == 0) {                       // no codes at all: any use of the table is an error
        for (int i = 0; i < first_size; ++i) table[i] = ENTRY_INVALID;
        return true;
    }
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;           // over-subscribed
    }
    if (left > 0 && (kind == T_PRECODE || max_len != 1)) return false;   // incomplete (inftrees.c rule)
    if (left > 0)                             // a complete code covers every first-level entry below
        for (int i = 0; i < first_size; ++i) table[i] = ENTRY_INVALID;
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
        code = (code + (uint32_t)count[l - 1] * (l > 1)) << 1;
        next_code[l] = code;
    }
    // canonical codes, symbol order within a length
    uint16_t sym_code[LITLEN_SYMS];
    for (int s = 0; s < n_syms; ++s)
        if (lens[s]) sym_code[s] = (uint16_t)bit_reverse(next_code[lens[s]]++, lens[s]);
    // second-level tables: one per first-level prefix that long codes share, as wide as its longest code
    int sub_bits[1 << LITLEN_BITS];
    bool any_long = max_len > first_bits;
    if (any_long) {
        for (int i = 0; i < first_size; ++i) sub_bits[i] = 0;
        for (int s = 0; s < n_syms; ++s)
            if (lens[s] > first_bits) {
                const int prefix = sym_code[s] & (first_size - 1);
                if (lens[s] - first_bits > sub_bits[prefix]) sub_bits[prefix] = lens[s] - first_bits;
            }
        int next_free = first_size;
        for (int i = 0; i < first_size; ++i)
            if (sub_bits[i]) {
                const int size = 1 << sub_bits[i];
                if (next_free + size > table_cap) return false;
                table[i] = K_POINTER | ((uint32_t)next_free << 16) | ((uint32_t)sub_bits[i] << 8) | (uint32_t)first_bits;
                for (int k = 0; k < size; ++k) table[next_free + k] = ENTRY_INVALID;
                next_free += size;
            }
    }
    for (int s = 0; s < n_syms; ++s) {
        const int len = lens[s];
        if (!len) continue;
        if (len <= first_bits) {
            const uint32_t e = symbol_entry(kind, s, len);
            for (int i = sym_code[s]; i < first_size; i += 1 << len) table[i] = e;
        } else {
            const int prefix = sym_code[s] & (first_size - 1);
            const uint32_t ptr = table[prefix];
            const int base = (int)(ptr >> 16), bits = (int)((ptr >> 8) & 15);
            const int rest = len - first_bits;
            // the entry drops only the bits after the first level; code length field keeps the
            // position of the extra bits relative to what is left in the buffer
            uint32_t e = symbol_entry(kind, s, rest);
            for (int i = sym_code[s] >> first_bits; i < (1 << bits); i += 1 << rest) table[base + i] = e;
        }
    }
    return true;
}

// Second pass over the first level of a literal/length table, after the distance table of the block exists.
//
// Pairs of literals: where the code of a literal leaves room in the index for the whole code of the
// literal after it, the entry delivers both (literal decode is a dependent load -> shift -> mask
// chain; two per lookup nearly halve it).  Ascending order: entry i >> len (the bits after the
// first code, zero-extended) has been visited already or is a plain single -- either way its first
// literal and the length of its first code are what is needed.
//
// Fused matches: base-call planes deflate into short chance matches (3..5 bytes) at distances of
// thousands: a short length code followed by a short distance code.  Where a length code without
// extra bits leaves room in the index for the whole distance code, one entry delivers the match --
// length, distance base and where the distance's extra bits are: one dependent table load per match
// instead of two.
int combine_entries(uint32_t *litlen, const uint32_t *dist) {
    const int first_size = 1 << LITLEN_BITS;
    int fused = 0;
    for (int i = 0; i < first_size; ++i) {
        const uint32_t e1 = litlen[i];
        const uint32_t k1 = e1 & K_MASK;
        if (k1 == K_LITERAL) {
            const int l1 = (int)NB(e1);
            if (l1 >= LITLEN_BITS) continue;
            const uint32_t e2 = litlen[i >> l1];
            if ((e2 & K_MASK) != K_LITERAL) continue;
            const int l2 = (int)((e2 >> 8) & 15);
            if (l1 + l2 > LITLEN_BITS) continue;
            litlen[i] = K_LITERAL | (2u << 12) | (((e2 >> 16) & 0xffu) << 24) | (e1 & 0x00ff0000u) | ((uint32_t)l1 << 8) | (uint32_t)(l1 + l2);
        } else if (k1 == K_BASE) {
            const int l1 = (int)((e1 >> 8) & 15);
            if ((int)NB(e1) != l1) continue;                          // the length has extra bits
            const uint32_t length = e1 >> 16;
            if (length > 10 || l1 >= LITLEN_BITS) continue;
            const uint32_t e2 = dist[(i >> l1) & ((1 << DIST_BITS) - 1)];
            if ((e2 & K_MASK) != K_BASE) continue;
            const int l2 = (int)((e2 >> 8) & 15);
            if (l1 + l2 > LITLEN_BITS) continue;                      // the index does not hold the whole distance code
            const int extra = (int)NB(e2) - l2;
            litlen[i] = K_MATCH | (e2 & 0xffff0000u) | (length << 12) | ((uint32_t)(l1 + l2) << 8) | (uint32_t)(l1 + l2 + extra);
            ++fused;
        }
    }
    return fused;
}

struct Inflater {
    uint32_t litlen[LITLEN_TABLE];
    uint32_t dist[DIST_TABLE];
    uint32_t fixed_litlen[LITLEN_TABLE];
    uint32_t fixed_dist[DIST_TABLE];
    uint32_t pre[1 << PRE_BITS];
    bool fixed_ready = false, fixed_fused = false;
    const char *why = "";

    int fail(int rc, const char *msg) {
        why = msg;
        return rc;
    }

    // One raw deflate stream: in[0, in_len) -> out[0, out_cap).  *in_used / *out_len report
    // what was consumed / produced (also on error, as far as it got).
    int inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *in_used, size_t *out_len);
};

inline void store16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }

inline uint64_t load64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}

// Two builds of the decoder, picked once at load time (ifunc): with BMI2 the variable shifts of the
// bit buffer are SHRX / SHLX (no shuffling through CL, no flag dependency) -- 15-20 % on the
// branch-free step, whose cost is its instruction count.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(WD_INFLATE_NO_MULTIVERSION)
__attribute__((target_clones("default", "bmi2")))
#endif
int Inflater::inflate_raw(const uint8_t *const in, const size_t in_len, uint8_t *const out, const size_t out_cap,
                          size_t *in_used, size_t *out_len) {
    const uint8_t *ip = in;
    const uint8_t *const in_end = in + in_len;
    uint8_t *op = out;
    uint8_t *const out_end = out + out_cap;
    uint64_t bitbuf = 0;
    int bitcnt = 0;          // valid bits in bitbuf; may go negative once the input is exhausted
    int rc = INF_OK;
    bool last = false;

// byte-wise refill for headers and the careful loop: never reads past in_end; once the input is
// exhausted the buffer is padded with zeros and `bitcnt` keeps honest books (negative = over-read)
#define REFILL_SAFE()                                              \
    do {                                                           \
        while (bitcnt < 56 && ip < in_end) {                        \
            bitbuf |= (uint64_t)*ip++ << bitcnt;                   \
            bitcnt += 8;                                           \
        }                                                          \
    } while (0)
#define DROP(n) (bitbuf >>= (n), bitcnt -= (int)(n))
#define BITS(n) ((uint32_t)bitbuf & ((1u << (n)) - 1u))

    do {
        REFILL_SAFE();
        if (bitcnt < 3) { rc = fail(INF_TRUNCATED, "stream ends inside a block header"); goto done; }
        last = BITS(1);
        const uint32_t type = (BITS(3) >> 1);
        DROP(3);
        const uint32_t *lt, *dt;
        bool uniform = false;        // the block's table holds fused matches: take literals and matches through one branch-free step
        if (type == 0) {
            // stored: back to a byte boundary, hand unread whole bytes back to the input
            DROP(bitcnt & 7);
            ip -= bitcnt >> 3;
            bitbuf = 0;
            bitcnt = 0;
            if (in_end - ip < 4) { rc = fail(INF_TRUNCATED, "stream ends inside a stored block header"); goto done; }
            const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ nlen) != 0xffff) { rc = fail(INF_BAD_DATA, "invalid stored block lengths"); goto done; }
            if ((size_t)(in_end - ip) < len) {
                // deliver what is there (zlib would too) so that out_len is meaningful, then report
                const size_t have = (size_t)(in_end - ip);
                const size_t room = (size_t)(out_end - op);
                const size_t n = have < room ? have : room;
                memcpy(op, ip, n);
                op += n;
                ip += n;
                rc = fail(INF_TRUNCATED, "stream ends inside a stored block");
                goto done;
            }
            if ((size_t)(out_end - op) < len) {
                memcpy(op, ip, (size_t)(out_end - op));      // fill what fits, like the Huffman blocks do
                ip += out_end - op;
                op = out_end;
                rc = fail(INF_OUT_FULL, "output buffer full");
                goto done;
            }
            memcpy(op, ip, len);
            op += len;
            ip += len;
            continue;
        } else if (type == 1) {
            if (!fixed_ready) {
                uint8_t lens[LITLEN_SYMS];
                for (int s = 0; s < 144; ++s) lens[s] = 8;
                for (int s = 144; s < 256; ++s) lens[s] = 9;
                for (int s = 256; s < 280; ++s) lens[s] = 7;
                for (int s = 280; s < 288; ++s) lens[s] = 8;
                build_table(T_LITLEN, lens, 288, LITLEN_BITS, fixed_litlen, LITLEN_TABLE);
                for (int s = 0; s < 32; ++s) lens[s] = 5;
                build_table(T_DIST, lens, 32, DIST_BITS, fixed_dist, DIST_TABLE);
                fixed_fused = combine_entries(fixed_litlen, fixed_dist) > 0;
                fixed_ready = true;
            }
            lt = fixed_litlen;
            dt = fixed_dist;
            uniform = fixed_fused;
        } else if (type == 2) {
            REFILL_SAFE();
            if (bitcnt < 14) { rc = fail(INF_TRUNCATED, "stream ends inside a block header"); goto done; }
            const int n_lit = (int)BITS(5) + 257;
            DROP(5);
            const int n_dist = (int)BITS(5) + 1;
            DROP(5);
            const int n_pre = (int)BITS(4) + 4;
            DROP(4);
            if (n_lit > 286 || n_dist > 30) { rc = fail(INF_BAD_DATA, "too many length or distance symbols"); goto done; }
            uint8_t pre_lens[19] = {0};
            for (int i = 0; i < n_pre; ++i) {
                REFILL_SAFE();
                if (bitcnt < 3) { rc = fail(INF_TRUNCATED, "stream ends inside a block header"); goto done; }
                pre_lens[PRE_ORDER[i]] = (uint8_t)BITS(3);
                DROP(3);
            }
            if (!build_table(T_PRECODE, pre_lens, 19, PRE_BITS, pre, 1 << PRE_BITS)) {
                rc = fail(INF_BAD_DATA, "invalid code lengths set");
                goto done;
            }
            uint8_t lens[LITLEN_SYMS + DIST_SYMS + 138];
            int have = 0;
            const int total = n_lit + n_dist;
            while (have < total) {
                REFILL_SAFE();
                const uint32_t e = pre[BITS(PRE_BITS)];
                const int nb = (int)NB(e);
                if (nb > bitcnt) { rc = fail(INF_TRUNCATED, "stream ends inside a block header"); goto done; }
                if ((e & K_MASK) != K_LITERAL) { rc = fail(INF_BAD_DATA, "invalid code lengths set"); goto done; }
                const int sym = (int)(e >> 16);
                DROP(nb);
                if (bitcnt < (sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0)) {
                    rc = fail(INF_TRUNCATED, "stream ends inside a block header");
                    goto done;
                }
                if (sym < 16) {
                    lens[have++] = (uint8_t)sym;
                } else {
                    int rep;
                    uint8_t val = 0;
                    if (sym == 16) {
                        if (have == 0) { rc = fail(INF_BAD_DATA, "invalid bit length repeat"); goto done; }
                        val = lens[have - 1];
                        rep = 3 + (int)BITS(2);
                        DROP(2);
                    } else if (sym == 17) {
                        rep = 3 + (int)BITS(3);
                        DROP(3);
                    } else {
                        rep = 11 + (int)BITS(7);
                        DROP(7);
                    }
                    if (have + rep > total) { rc = fail(INF_BAD_DATA, "invalid bit length repeat"); goto done; }
                    memset(lens + have, val, (size_t)rep);
                    have += rep;
                }
            }
            if (lens[256] == 0) { rc = fail(INF_BAD_DATA, "invalid code -- missing end-of-block"); goto done; }
            uint8_t dl[DIST_SYMS] = {0};
            memcpy(dl, lens + n_lit, (size_t)n_dist);
            memset(lens + n_lit, 0, (size_t)(LITLEN_SYMS - n_lit));
            if (!build_table(T_LITLEN, lens, LITLEN_SYMS, LITLEN_BITS, litlen, LITLEN_TABLE)) {
                rc = fail(INF_BAD_DATA, "invalid literal/lengths set");
                goto done;
            }
            if (!build_table(T_DIST, dl, DIST_SYMS, DIST_BITS, dist, DIST_TABLE)) {
                rc = fail(INF_BAD_DATA, "invalid distances set");
                goto done;
            }
            uniform = combine_entries(litlen, dist) > 0;
            lt = litlen;
            dt = dist;
        } else {
            rc = fail(INF_BAD_DATA, "invalid block type");
            goto done;
        }

        // ---- symbols of one Huffman block -----------------------------------------------
        for (;;) {
            // Fast loop: while 16 input bytes and a longest match plus the copy overshoot fit,
            // nothing inside needs a bounds check.
            if (in_end - ip >= 16 && out_end - op >= 258 + 48) {
                const uint8_t *const in_fast = in_end - 16;
                uint8_t *const out_fast = out_end - (258 + 48);
                bool block_done = false;
                // whole-byte view of the buffer for the branch-free refill
                bitbuf &= (1ull << bitcnt) - 1ull;
// After a refill at least 56 bits are valid.  A refill only adds bits above the valid ones, so an
// entry looked up before it stays the right one: every iteration refills once and never looks the
// same symbol up twice, and the lookup of the next symbol is issued before the match copy.
#define REFILL_FAST()                          \
    do {                                       \
        bitbuf |= load64(ip) << bitcnt;        \
        ip += (63 - bitcnt) >> 3;              \
        bitcnt |= 56;                          \
    } while (0)
#define LOOKUP_LITLEN() lt[bitbuf & ((1u << LITLEN_BITS) - 1u)]
                REFILL_FAST();
                uint32_t e = LOOKUP_LITLEN();
// Blocks that mix literals and short matches (every base-call plane): which of the two comes next
// is a coin toss per symbol, and a mispredicted branch costs more than either.  Literal entries (one
// or two bytes) and fused match entries have one shape -- bits to drop, bytes produced, 16 payload
// bits -- so one step serves both without a branch: the payload is stored as two literal bytes, then
// 16 bytes are copied -- for a match from `op - distance` over them, for a literal from a block of
// zeros to just behind them (overwritten by the next symbol).  The step is bound by its instruction
// count (no store-to-load forwarding, no loop-carried memory dependency).  Only what is rare
// branches: a distance below 16 (the 16-byte copy would overlap) and, during the first 32 KB of
// output, a distance that reaches in front of it.
#define UNIFORM_STEP(FAR_OK)                                                                                        \
    {                                                                                                               \
        const uint32_t nb = e & 63u;             /* bits 5 and 6 are clear in both kinds */                         \
        const uint32_t pos = (e >> 8) & 15u;                                                                        \
        const uint32_t payload = e >> 16;                                                                           \
        const uint32_t distance = payload + (((uint32_t)bitbuf & ((1u << nb) - 1u)) >> pos);                        \
        const uintptr_t adv = (e >> 12) & 15u;                                                                      \
        const uintptr_t m = (uintptr_t)(intptr_t)((int32_t)(e << 24) >> 31);      /* all ones for a match */        \
        bitbuf >>= nb;                                                                                              \
        bitcnt -= (int)nb;                                                                                          \
        e = LOOKUP_LITLEN();                                                                                        \
        if (__builtin_expect(distance < 16 || (!(FAR_OK) && distance > (size_t)(op - out)), 0)) {                   \
            if (m) {                                                                                                \
                if (distance > (size_t)(op - out)) { rc = fail(INF_BAD_DATA, "invalid distance too far back"); goto done; } \
                for (uintptr_t k = 0; k < adv; ++k) op[k] = op[(ptrdiff_t)k - (ptrdiff_t)distance];                 \
            } else {                                                                                                \
                store16(op, (uint16_t)payload);                                                                     \
            }                                                                                                       \
        } else {                                                                                                    \
            store16(op, (uint16_t)payload);                                                                         \
            const uintptr_t su = (uintptr_t)scratch, ou = (uintptr_t)op;                                            \
            memcpy((uint8_t *)(ou + (adv & ~m)), (const uint8_t *)(su + ((ou - distance - su) & m)), 16);           \
        }                                                                                                           \
        op += adv;                                                                                                  \
    }
                alignas(16) static const uint8_t scratch[16] = {0};
                // every distance a stream can code reaches back at most 32768 bytes: beyond that much output only
                // distances below 16 need a second look
                const bool far_ok = op - out >= 32768;
                do {
                    if (uniform && (e & (3u << 5)) == 0) {          // K_LITERAL or K_MATCH
                        if (far_ok) {
                            UNIFORM_STEP(true);                     // <= 23 bits each: two fit one refill
                            if ((e & (3u << 5)) == 0) UNIFORM_STEP(true);
                        } else {
                            UNIFORM_STEP(false);
                            if ((e & (3u << 5)) == 0) UNIFORM_STEP(false);
                        }
                        REFILL_FAST();
                        continue;
                    }
                    if ((e & K_MASK) == K_LITERAL) {
                        // up to three first-level entries (one or two literals each, <= 3 x LITLEN_BITS bits) leave
                        // more than the 15 valid bits the lookup after them may need
                        DROP(NB(e));
                        store16(op, (uint16_t)(e >> 16));
                        op += (e >> 12) & 15;
                        e = LOOKUP_LITLEN();
                        if ((e & K_MASK) == K_LITERAL) {
                            DROP(NB(e));
                            store16(op, (uint16_t)(e >> 16));
                            op += (e >> 12) & 15;
                            e = LOOKUP_LITLEN();
                            if ((e & K_MASK) == K_LITERAL) {
                                DROP(NB(e));
                                store16(op, (uint16_t)(e >> 16));
                                op += (e >> 12) & 15;
                                e = LOOKUP_LITLEN();
                            }
                        }
                        REFILL_FAST();
                        continue;
                    }
                    uint32_t length, distance;
                    if ((e & K_MASK) == K_MATCH) {
                        // length and distance code from one entry; only the distance's extra bits are left to read
                        length = (e >> 12) & 15;
                        distance = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                        DROP(NB(e));
                    } else {
                        if ((e & K_MASK) == K_POINTER) {
                            DROP(LITLEN_BITS);
                            e = lt[(e >> 16) + BITS((e >> 8) & 15)];
                            if ((e & K_MASK) == K_LITERAL) {
                                DROP(NB(e));
                                *op++ = (uint8_t)(e >> 16);
                                e = LOOKUP_LITLEN();
                                REFILL_FAST();
                                continue;
                            }
                        }
                        if ((e & K_MASK) == K_SPECIAL) {
                            if ((e >> 16) != ENTRY_EOB_PAYLOAD) { rc = fail(INF_BAD_DATA, "invalid literal/length code"); goto done; }
                            DROP(NB(e));
                            block_done = true;
                            break;
                        }
                        // length (<= 20 bits with the pointer step), distance (<= 28 bits): 48 of the 56
                        length = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                        DROP(NB(e));
                        e = dt[bitbuf & ((1u << DIST_BITS) - 1u)];
                        if ((e & K_MASK) == K_POINTER) {
                            DROP(DIST_BITS);
                            e = dt[(e >> 16) + BITS((e >> 8) & 15)];
                        }
                        if ((e & K_MASK) != K_BASE) { rc = fail(INF_BAD_DATA, "invalid distance code"); goto done; }
                        distance = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                        DROP(NB(e));
                    }
                    if (distance > (size_t)(op - out)) { rc = fail(INF_BAD_DATA, "invalid distance too far back"); goto done; }
                    REFILL_FAST();
                    e = LOOKUP_LITLEN();
                    const uint8_t *src = op - distance;
                    uint8_t *dst = op;
                    op += length;
                    if (distance >= 16) {
                        memcpy(dst, src, 16);
                        if (length > 16) {
                            memcpy(dst + 16, src + 16, 16);
                            if (length > 32) {
                                dst += 32;
                                src += 32;
                                do {
                                    memcpy(dst, src, 16);
                                    dst += 16;
                                    src += 16;
                                } while (dst < op);
                            }
                        }
                    } else if (distance == 1) {
                        memset(dst, *src, length);
                    } else {
                        // short period: byte by byte keeps the overlap semantics
                        for (uint32_t k = 0; k < length; ++k) dst[k] = src[k];
                    }
                } while (ip <= in_fast && op <= out_fast);
#undef REFILL_FAST
#undef LOOKUP_LITLEN
#undef UNIFORM_STEP
                if (block_done) break;
                continue;       // re-evaluate: the careful loop below takes over near the ends
            }

            // Careful loop: one symbol at a time with every bound checked.
            REFILL_SAFE();
            uint32_t e = lt[BITS(LITLEN_BITS)];
            if ((e & K_MASK) == K_POINTER) {
                DROP(LITLEN_BITS);
                e = lt[(e >> 16) + BITS((e >> 8) & 15)];
            }
            if ((e & K_MASK) == K_LITERAL) {
                DROP((e >> 8) & 15);         // one literal at a time here: the first code of the entry
                if (bitcnt < 0) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
                if (op == out_end) { rc = fail(INF_OUT_FULL, "output buffer full"); goto done; }
                *op++ = (uint8_t)(e >> 16);
                continue;
            }
            if ((e & K_MASK) == K_SPECIAL) {
                if ((int)NB(e) > bitcnt && ip == in_end) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
                if ((e >> 16) != ENTRY_EOB_PAYLOAD) { rc = fail(INF_BAD_DATA, "invalid literal/length code"); goto done; }
                DROP(NB(e));
                break;
            }
            uint32_t length, distance;
            if ((e & K_MASK) == K_MATCH) {
                length = (e >> 12) & 15;
                distance = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                DROP(NB(e));
                if (bitcnt < 0) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
            } else {
                length = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                DROP(NB(e));
                if (bitcnt < 0) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
                REFILL_SAFE();
                e = dt[BITS(DIST_BITS)];
                if ((e & K_MASK) == K_POINTER) {
                    DROP(DIST_BITS);
                    e = dt[(e >> 16) + BITS((e >> 8) & 15)];
                }
                if ((e & K_MASK) != K_BASE) {
                    if ((int)NB(e) > bitcnt && ip == in_end) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
                    rc = fail(INF_BAD_DATA, "invalid distance code");
                    goto done;
                }
                distance = (e >> 16) + (((uint32_t)bitbuf >> ((e >> 8) & 15)) & ((1u << (NB(e) - ((e >> 8) & 15))) - 1u));
                DROP(NB(e));
                if (bitcnt < 0) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
            }
            if (distance > (size_t)(op - out)) { rc = fail(INF_BAD_DATA, "invalid distance too far back"); goto done; }
            if (length > (size_t)(out_end - op)) {
                // fill what fits, like zlib, then report
                while (op < out_end) { *op = *(op - distance); ++op; }
                rc = fail(INF_OUT_FULL, "output buffer full");
                goto done;
            }
            for (uint32_t k = 0; k < length; ++k) op[k] = op[(ptrdiff_t)k - (ptrdiff_t)distance];
            op += length;
        }
        if (bitcnt < 0) { rc = fail(INF_TRUNCATED, "stream ends inside a block"); goto done; }
    } while (!last);

done:
    // unread whole bytes go back to the caller (the gzip trailer follows on a byte boundary)
    if (bitcnt > 0) {
        size_t back = (size_t)(bitcnt >> 3);
        if (back > (size_t)(ip - in)) back = (size_t)(ip - in);
        ip -= back;
    }
    *in_used = (size_t)(ip - in);
    *out_len = (size_t)(op - out);
    return rc;
#undef REFILL_SAFE
#undef DROP
#undef BITS
}

// ---------------------------------------------------------------------------------------------
// gzip framing (RFC 1952) with gzip.open()'s reading rules (Lib/gzip.py: _GzipReader)
// ---------------------------------------------------------------------------------------------
struct GunzipResult {
    int status = WD_OK;
    size_t out_len = 0;
    std::string message;
};

void set_result(GunzipResult &r, int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    r.status = status;
    r.message = buf;
}

// All members of in[0, n) -> out[0, cap).
void gunzip_members(Inflater &inf, const uint8_t *in, size_t n, uint8_t *out, size_t cap, GunzipResult &res) {
    size_t pos = 0, produced = 0;
    bool first = true;
    for (;;) {
        if (!first) {
            while (pos < n && in[pos] == 0) ++pos;      // zero padding between / after members
            if (pos == n) break;
        }
        if (pos == n) break;                            // an empty file reads as b"" (gzip.py)
        if (n - pos < 2) { set_result(res, WD_E_EOF, "Compressed file ended before the end-of-stream marker was reached"); break; }
        if (in[pos] != 0x1f || in[pos + 1] != 0x8b) {
            set_result(res, WD_E_DATA, "Not a gzipped file (%02x %02x)", in[pos], in[pos + 1]);
            break;
        }
        if (n - pos < 10) { set_result(res, WD_E_EOF, "Compressed file ended before the end-of-stream marker was reached"); break; }
        if (in[pos + 2] != 8) { set_result(res, WD_E_DATA, "Unknown compression method"); break; }
        const uint8_t flg = in[pos + 3];
        size_t p = pos + 10;
        bool short_header = false;
        if (flg & 4) {                                   // FEXTRA
            if (n - p < 2) short_header = true;
            else {
                const size_t xlen = in[p] | (in[p + 1] << 8);
                p += 2;
                if (n - p < xlen) short_header = true;
                else p += xlen;
            }
        }
        for (int field = 0; field < 2 && !short_header; ++field)      // FNAME, FCOMMENT
            if (flg & (field == 0 ? 8 : 16)) {
                while (p < n && in[p] != 0) ++p;
                if (p == n) short_header = true;
                else ++p;
            }
        if (!short_header && (flg & 2)) {                // FHCRC (read, not verified: as gzip.py)
            if (n - p < 2) short_header = true;
            else p += 2;
        }
        if (short_header) { set_result(res, WD_E_EOF, "Compressed file ended before the end-of-stream marker was reached"); break; }
        size_t used = 0, got = 0;
        const int rc = inf.inflate_raw(in + p, n - p, out + produced, cap - produced, &used, &got);
        const uint8_t *member_out = out + produced;
        produced += got;
        if (rc == INF_TRUNCATED) { set_result(res, WD_E_EOF, "Compressed file ended before the end-of-stream marker was reached"); break; }
        if (rc == INF_BAD_DATA) { set_result(res, WD_E_DATA, "Error -3 while decompressing data: %s", inf.why); break; }
        if (rc == INF_OUT_FULL) { set_result(res, WD_E_CAPACITY, "inflated data exceed the %zu bytes the caller expects", cap); break; }
        p += used;
        if (n - p < 8) { set_result(res, WD_E_EOF, "Compressed file ended before the end-of-stream marker was reached"); break; }
        const uint32_t want_crc = in[p] | (in[p + 1] << 8) | (in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
        const uint32_t want_len = in[p + 4] | (in[p + 5] << 8) | (in[p + 6] << 16) | ((uint32_t)in[p + 7] << 24);
        const uint32_t crc = crc32_update(0, member_out, got);
        if (crc != want_crc) { set_result(res, WD_E_DATA, "CRC check failed 0x%x != 0x%x", want_crc, crc); break; }
        if (want_len != (uint32_t)(got & 0xffffffffu)) { set_result(res, WD_E_DATA, "Incorrect length of data produced"); break; }
        pos = p + 8;
        first = false;
    }
    res.out_len = produced;
}

bool read_whole(const char *path, uint64_t offset, uint64_t size, std::vector<uint8_t> &buf, GunzipResult &res) {
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        set_result(res, errno == ENOENT || errno == ENOTDIR ? WD_E_NOENT : WD_E_IO, "[Errno %d] %s: '%s'", errno, strerror(errno), path);
        return false;
    }
    if (size == 0) {
        struct stat st;
        if (fstat(fd, &st) != 0) {
            set_result(res, WD_E_IO, "[Errno %d] %s: '%s'", errno, strerror(errno), path);
            close(fd);
            return false;
        }
        size = (uint64_t)st.st_size > offset ? (uint64_t)st.st_size - offset : 0;
    }
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd, (off_t)offset, (off_t)size, POSIX_FADV_SEQUENTIAL);
#endif
    buf.resize(size);
    uint64_t got = 0;
    while (got < size) {
        const ssize_t k = pread(fd, buf.data() + got, size - got, (off_t)(offset + got));
        if (k < 0) {
            if (errno == EINTR) continue;
            set_result(res, WD_E_IO, "[Errno %d] %s: '%s'", errno, strerror(errno), path);
            close(fd);
            return false;
        }
        if (k == 0) break;      // shorter than announced: the gzip layer reports the truncation
        got += (uint64_t)k;
    }
    close(fd);
    buf.resize(got);
    return true;
}

void run_job(Inflater &inf, std::vector<uint8_t> &scratch, wd_inflate_job &job) {
    GunzipResult res;
    const uint8_t *src = job.src;
    size_t n = (size_t)job.size;
    if (job.path != nullptr) {
        if (!read_whole(job.path, job.offset, job.size, scratch, res)) {
            job.status = res.status;
            job.out_len = 0;
            snprintf(job.message, sizeof job.message, "%s", res.message.c_str());
            return;
        }
        src = scratch.data();
        n = scratch.size();
    } else if (src == nullptr) {
        job.status = WD_E_ARG;
        job.out_len = 0;
        snprintf(job.message, sizeof job.message, "job has neither a path nor a source buffer");
        return;
    }
    if (job.dst == nullptr && job.dst_cap != 0) {
        job.status = WD_E_ARG;
        job.out_len = 0;
        snprintf(job.message, sizeof job.message, "job has no destination");
        return;
    }
    gunzip_members(inf, src, n, job.dst, (size_t)job.dst_cap, res);
    job.status = res.status;
    job.out_len = res.out_len;
    snprintf(job.message, sizeof job.message, "%s", res.message.c_str());
}

}   // namespace

extern "C" {

uint32_t wd_crc32(uint32_t crc, const uint8_t *data, size_t n) {
    return (data == nullptr || n == 0) ? crc : crc32_update(crc, data, n);
}

int wd_inflate_batch(wd_inflate_job *jobs, size_t n_jobs, int threads) {
    if (jobs == nullptr && n_jobs != 0) {
        wd::set_error("wd_inflate_batch: null job list");
        return WD_E_ARG;
    }
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if ((size_t)threads > n_jobs) threads = (int)n_jobs;
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        Inflater *inf = new Inflater;        // ~30 KB of tables: off the thread's stack
        std::vector<uint8_t> scratch;
        for (;;) {
            const size_t k = next.fetch_add(1, std::memory_order_relaxed);
            if (k >= n_jobs) break;
            run_job(*inf, scratch, jobs[k]);
        }
        delete inf;
    };
    if (threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        pool.reserve((size_t)threads - 1);
        for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
    }
    for (size_t k = 0; k < n_jobs; ++k)
        if (jobs[k].status != WD_OK) {
            wd::set_error("%s", jobs[k].message);
            return jobs[k].status;
        }
    return WD_OK;
}

int wd_gunzip(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    wd_inflate_job job;
    memset(&job, 0, sizeof job);
    job.src = src;
    job.size = n;
    job.dst = dst;
    job.dst_cap = cap;
    if (src == nullptr) {
        wd::set_error("wd_gunzip: null source");
        return WD_E_ARG;
    }
    const int rc = wd_inflate_batch(&job, 1, 1);
    if (out_len != nullptr) *out_len = (size_t)job.out_len;
    return rc;
}

}   // extern "C"
