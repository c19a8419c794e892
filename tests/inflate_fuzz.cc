#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../include/welldup.h"
// Sanitizer harness for csrc/wd_inflate.cc (built by tests/test_inflate_staging.py with
// -fsanitize=address,undefined): valid gzip files given on the command line are truncated and
// bit-flipped, source and destination live in exact-size heap blocks, so any read or write outside
// them -- or any undefined shift / overflow on the way -- aborts the run.
#include <cstdarg>
namespace wd {
void set_error(const char *, ...) {}
}
int main(int argc, char **argv) {
    std::mt19937_64 rng(7);
    long runs = 0, ok = 0;
    for (int a = 1; a < argc; ++a) {
        FILE *f = fopen(argv[a], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> in(n); if (fread(in.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
        if (n > 300000) { in.resize(300000); n = 300000; }
        for (int it = 0; it < 400; ++it) {
            // exact-size heap copies so ASAN sees any over-read / over-write
            size_t len = it % 7 == 0 ? rng() % n + 1 : n;
            unsigned char *src = (unsigned char *)malloc(len); memcpy(src, in.data(), len);
            int flips = it % 5;
            for (int k = 0; k < flips; ++k) src[10 + rng() % (len > 10 ? len - 10 : 1) % len] ^= 1u << (rng() % 8);
            size_t cap = it % 3 == 0 ? rng() % 100000 : 1500000;
            unsigned char *dst = (unsigned char *)malloc(cap ? cap : 1);
            size_t got = 0;
            int rc = wd_gunzip(src, len, dst, cap, &got);
            if (got > cap) { printf("overrun!\n"); return 1; }
            // an untouched stream with room for its output must inflate, and its CRC-32 (checked inside) must match
            if (flips == 0 && len == (size_t)n && cap == 1500000 && rc != 0) { printf("valid stream refused: %d\n", rc); return 1; }
            ++runs; ok += rc == 0;
            free(src); free(dst);
        }
    }
    printf("fuzz runs %ld, clean %ld\n", runs, ok);
}
