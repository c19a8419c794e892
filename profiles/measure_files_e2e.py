"""A lane counted FROM COMPRESSED FILES: run folder on disk -> report counters.

bench.py's `value` and `e2e` start from inflated planes (BASELINE.json north_star
keeps gunzip outside the kernel roofline); this script times what a user waits
for: `.filter` + `.bcl.gz` files of a HiSeq-4000-shaped lane (4 309 650 wells,
2500 targets x 5 rings, 50 compared cycles) through staging.lane_batches
(native inflate threads -> page-locked planes, next batch inflating while the
GPU counts) and wd_count on host-mapped tiles.  D distinct tiles are written
(gzip level 1, like bench.py's host_inflate sample) and linked under T tile
names; the page cache is warm.  Beside it: the same files through zlib
(gzip.open().read(), one thread per file as the old Python staging did) for
the inflate alone.  One JSON line on stdout.

    python profiles/measure_files_e2e.py [--tiles 32] [--distinct 8] [--dir /tmp/wd_run]
"""
import argparse
import gzip
import json
import os
import random
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_WELLS, ROW_LEN, N_CYCLES, N_TARGETS, LEVELS = 4309650, 1571, 50, 2500, 5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=32)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--dir", default="/tmp/wd_run")
    ap.add_argument("--wells", type=int, default=N_WELLS)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--dry", action="store_true", help="no GPU: files and inflate only (pageable blocks)")
    args = ap.parse_args()
    from well_duplicates_b200 import staging, synth
    from well_duplicates_b200.engine import Engine
    from well_duplicates_b200.reader import BCLReader

    n = args.wells
    ldir = synth.basecalls_dir(args.dir, 1)
    os.makedirs(ldir, exist_ok=True)
    names = ["%d%d%02d" % (s, w, t) for s in (1, 2) for w in (1, 2) for t in range(1, 25)][:args.tiles]
    t0 = time.perf_counter()
    comp_bytes = 0

    def write_plane(job):
        d, c, plane = job
        path = os.path.join(ldir, "C%d.1" % (c + 1), "d%d.bcl.gz" % d)
        with gzip.open(path, "wb", compresslevel=1) as fh:
            fh.write(synth.bcl_plane_bytes(plane))
        return os.path.getsize(path)

    for c in range(N_CYCLES):
        os.makedirs(os.path.join(ldir, "C%d.1" % (c + 1)), exist_ok=True)
    with ThreadPoolExecutor(max_workers=args.threads) as pool:
        for d in range(args.distinct):
            td = synth.make_tile_fast(1000 + d, n, N_CYCLES, ROW_LEN)
            with open(os.path.join(ldir, "d%d.filter" % d), "wb") as fh:
                fh.write(synth.filter_file_bytes(td.filt))
            comp_bytes += sum(pool.map(write_plane, [(d, c, td.planes[c]) for c in range(N_CYCLES)]))
    for k, name in enumerate(names):
        d = k % args.distinct
        for src, dst in [("d%d.filter" % d, "s_1_%s.filter" % name)] + \
                        [(os.path.join("C%d.1" % (c + 1), "d%d.bcl.gz" % d), os.path.join("C%d.1" % (c + 1), "s_1_%s.bcl.gz" % name))
                         for c in range(N_CYCLES)]:
            p = os.path.join(ldir, dst)
            if os.path.lexists(p):
                os.remove(p)
            os.link(os.path.join(ldir, src), p)
    # the scratch files d*.filter would match a tile-name regex only by accident; keep them out of the way
    for d in range(args.distinct):
        os.rename(os.path.join(ldir, "d%d.filter" % d), os.path.join(ldir, "d%d.filter.src" % d))
    t_write = time.perf_counter() - t0

    eng = None if args.dry else Engine(0)
    X, Y = synth.hex_lattice(n, ROW_LEN)
    random.seed(13)
    centres = np.array(random.sample(range(n), N_TARGETS), dtype=np.uint32)
    if eng is not None:
        eng.load_locs(synth.xy_to_locs_floats(X, Y))
        offs, idx = eng.ring_query(centres, LEVELS)
        eng.load_targets(centres, offs, idx, LEVELS)
    rd = BCLReader(args.dir, engine=eng)
    wanted = list(range(N_CYCLES))
    st = staging.Stager(threads=args.threads, pinned=not args.dry)

    def lane():
        rows, gpu_s, inflated, batches = [], 0.0, 0, 0
        t0 = time.perf_counter()
        for got, batch in staging.lane_batches(st, lambda t: rd.get_tile(1, t), names, wanted):
            g0 = time.perf_counter()
            if eng is None:
                cnt = np.zeros((len(got), 1 + 5 * LEVELS), np.int64)
            else:
                plane_of = st.deliver(eng, batch, first_slot=0, zero_copy=True)
                _, cnt = eng.count(0, len(got), [plane_of[c] for c in wanted], 2, False, mode=0, per_target=False)
            gpu_s += time.perf_counter() - g0
            rows.append(cnt)
            inflated += batch.inflated_bytes
            batches += 1
        return time.perf_counter() - t0, gpu_s, inflated, batches, np.concatenate(rows)

    lane()                                                     # page-locked blocks allocated, page cache warm
    wall, gpu_s, inflated, batches, rows = lane()
    # inflate alone through zlib, as the reference's reader (and the old Python staging) runs it
    files = [os.path.join(ldir, "C%d.1" % (c + 1), "s_1_%s.bcl.gz" % t) for t in names[:max(1, args.tiles // 4)] for c in range(N_CYCLES)]

    def zread(path):
        with gzip.open(path, "rb") as fh:
            return len(fh.read())
    z0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=args.threads) as pool:
        zbytes = sum(pool.map(zread, files))
    zdt = time.perf_counter() - z0
    # the native inflate alone (no GPU work) for the same files
    i0 = time.perf_counter()
    tiles = [rd.get_tile(1, t) for t in names[:max(1, args.tiles // 4)]]
    ibytes = 0
    per = st.tiles_per_batch(n, N_CYCLES)
    for k in range(0, len(tiles), per):
        ibytes += st.load(tiles[k:k + per], wanted, 0).inflated_bytes
    idt = time.perf_counter() - i0
    out = {"what": "one lane from .filter/.bcl.gz files on disk to counters (page cache warm)", "tiles": len(names),
           "distinct_tiles": args.distinct, "wells_per_tile": n, "cycles": N_CYCLES, "targets_per_tile": N_TARGETS,
           "host_threads": args.threads, "compressed_bytes_per_tile": comp_bytes // args.distinct,
           "wall_s": wall, "targets_per_s": len(names) * N_TARGETS / wall, "inflated_gb_per_s": inflated / wall / 1e9,
           "gpu_side_s": gpu_s, "batches": batches, "lane96_seconds_at_this_rate": wall * 96 / len(names),
           "native_inflate_only": {"gb_per_s": ibytes / idt / 1e9, "seconds": idt, "tiles": len(tiles)},
           "zlib_gzip_open_only": {"gb_per_s": zbytes / zdt / 1e9, "seconds": zdt, "tiles": len(tiles)},
           "valid_targets_total": int(rows[:, 0].sum()), "dups_level1_total": int(rows[:, 2].sum()),
           "write_files_s": t_write}
    print(json.dumps(out))
    st.close()


if __name__ == "__main__":
    main()
