// ThreadSanitizer harness for wd_inflate_batch (built by tests/test_inflate_staging.py with
// -fsanitize=thread): 64 jobs over the files given on the command line, 8 native threads.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../include/welldup.h"
namespace wd {
void set_error(const char *, ...) {}
}
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::vector<wd_inflate_job> jobs(64);
    std::vector<std::vector<unsigned char>> outs(64, std::vector<unsigned char>(1 << 20));
    for (int k = 0; k < 64; ++k) {
        memset(&jobs[k], 0, sizeof(wd_inflate_job));
        jobs[k].path = argv[1 + k % (argc - 1)];
        jobs[k].dst = outs[k].data();
        jobs[k].dst_cap = outs[k].size();
    }
    const int rc = wd_inflate_batch(jobs.data(), jobs.size(), 8);
    size_t total = 0;
    for (auto &j : jobs) total += j.out_len;
    printf("batch rc %d total %zu\n", rc, total);
    return rc == 0 ? 0 : 1;
}
