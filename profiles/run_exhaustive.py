#!/usr/bin/env python3
"""One full-size tile through wd_count_exhaustive (driver for ncu captures)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from well_duplicates_b200 import synth  # noqa: E402
from well_duplicates_b200.engine import Engine  # noqa: E402

N, ROW, NCYC = synth.HISEQ4000_WELLS, synth.HISEQ4000_ROW_LEN, 50
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ham = len(sys.argv) > 2 and sys.argv[2] == "hamming"
eng = Engine(0)
X, Y = synth.hex_lattice(N, ROW)
eng.load_locs(synth.xy_to_locs_floats(X, Y))
td = synth.make_tile_fast(20261018, N, NCYC, ROW)
eng.tile_begin(0, N, NCYC)
eng.tile_put_filter(0, td.filt)
for c in range(NCYC):
    eng.tile_put_bcl(0, c, td.planes[c])
for _ in range(reps):
    t0 = time.perf_counter()
    cnt = eng.count_exhaustive(0, list(range(NCYC)), 5, 2, ham)
    print("exhaustive: %.3f ms, targets %d, dups %s" % (1e3 * (time.perf_counter() - t0), cnt[0], cnt[2::5].tolist()))
