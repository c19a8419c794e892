#!/usr/bin/env python3
"""Timings of the BASELINE.json configs that bench.py does not carry as its
headline (configs[1] is bench.py's):

  config 3  exhaustive mode, one full-size tile, every well a target
  config 4  stage 1: .locs of a full tile -> rings of 2500 sampled targets
  config 5  NovaSeq-style CBCL lane (4-bit planes, first cycles with every well,
            later ones with pass-filter wells only)

    python profiles/measure_configs.py [--configs 3,4,5] [--cbcl-tiles 352] > profiles/rNN_configs.jsonl

One JSON line per config.  GPU times are CUDA-event / synchronous-call times
with inputs resident in HBM; the CPU column is the C restatement of the
reference (oracle/) on a bounded sample, scaled linearly and labelled as such.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from well_duplicates_b200 import synth  # noqa: E402
from well_duplicates_b200.engine import Engine  # noqa: E402

N = synth.HISEQ4000_WELLS
ROW = synth.HISEQ4000_ROW_LEN
NCYC = 50
PEAK = 6557.8


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"])
    return PEAK


def sync_time(eng, fn, reps):
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    eng.sync()
    return (time.perf_counter() - t0) / reps


def config4(eng, args):
    from oracle import c_port as CP
    X, Y = synth.hex_lattice(N, ROW)
    xy = synth.xy_to_locs_floats(X, Y)
    random.seed(13)
    centres = np.array(random.sample(range(N), 2500), dtype=np.uint32)
    eng.load_locs(xy)
    eng.ring_query(centres, 5)
    t_load = sync_time(eng, lambda: eng.load_locs(xy), 5)
    t_query = sync_time(eng, lambda: eng.ring_query(centres, 5), 10)
    offs, idx = eng.ring_query(centres, 5)
    # CPU: the reference's scan (C port) for a sample of the same centres
    k = 100
    t0 = time.perf_counter()
    woffs, widx = CP.rings_csr(X, Y, centres[:k])
    t_cpu = (time.perf_counter() - t0) * 2500 / k
    ok = bool(np.array_equal(offs[: woffs.size], woffs) and np.array_equal(idx[: widx.size], widx))
    # algorithmic bytes: locs in, pixels out, cell key + record, CSR out
    alg = N * 8 + N * 8 + N * 16 + int(idx.size) * 4
    return {"config": "4: prepare_cluster_indexes neighbourhood build, full-tile .locs (%d wells), 2500 targets x 5 rings" % N,
            "locs_load_ms": 1e3 * t_load, "ring_query_ms": 1e3 * t_query, "total_ms": 1e3 * (t_load + t_query),
            "targets_per_s": 2500 / (t_load + t_query),
            "note": "locs_load = pageable H2D of 34.5 MB + K0 + K1 (grid build, once per flowcell); ring_query = K2 "
                    "count + scan + fill + D2H of the CSR; wall clock around the synchronous C-ABI calls",
            "algorithmic_bytes": alg, "achieved_gb_per_s": alg / (t_load + t_query) / 1e9, "hbm_peak_gb_per_s": peak(),
            "cpu_port_s_scaled": t_cpu, "cpu_sample": "%d of the 2500 centres, C restatement of get_indexes, 1 thread" % k,
            "matches_oracle_on_sample": ok, "reference_python_s": 180.9,
            "reference_python_note": "unmodified prepare_cluster_indexes.py -n 2500, measured in the authoring container (SURVEY 6)"}


def config3(eng, args):
    from oracle import c_port as CP
    X, Y = synth.hex_lattice(N, ROW)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    td = synth.make_tile_fast(20261018, N, NCYC, ROW)
    eng.tile_begin(0, N, NCYC)
    eng.tile_put_filter(0, td.filt)
    for c in range(NCYC):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(NCYC))
    out = {}
    for name, ham in (("levenshtein", False), ("hamming", True)):
        cnt = eng.count_exhaustive(0, order, 5, 2, ham)
        t = sync_time(eng, lambda: eng.count_exhaustive(0, order, 5, 2, ham), 3)
        out[name] = {"ms": 1e3 * t, "targets_per_s": N / t, "wells_compared_per_s": int(cnt[1::5].sum()) / t,
                     "valid_targets": int(cnt[0]), "wells_compared": int(cnt[1::5].sum()), "dups": cnt[2::5].tolist()}
    # CPU: cropped tile (first rows), every well a target, scaled by wells
    rows = 60
    n_c = rows * ROW
    Xc, Yc = X[:n_c], Y[:n_c]
    t0 = time.perf_counter()
    want = CP.count_exhaustive(Xc, Yc, [td.planes[c][:n_c] for c in order], ["bcl"] * NCYC, td.filt[:n_c], 5, 2, False)
    t_cpu = time.perf_counter() - t0
    eng.load_locs(synth.xy_to_locs_floats(Xc, Yc))
    eng.tile_begin(1, n_c, NCYC)
    eng.tile_put_filter(1, td.filt[:n_c])
    for c in range(NCYC):
        eng.tile_put_bcl(1, c, td.planes[c][:n_c])
    got = eng.count_exhaustive(1, order, 5, 2, False)
    alg = N * NCYC + N * 8 + N + 2 * N * 24
    t = out["levenshtein"]["ms"] / 1e3
    return {"config": "3: exhaustive mode, one full-size tile (%d wells), every well a target out to ring 5, 50 cycles" % N,
            **out, "algorithmic_bytes": alg, "achieved_gb_per_s": alg / t / 1e9, "hbm_peak_gb_per_s": peak(),
            "bound_note": "whole wd_count_exhaustive call (dense pack, prefix layout, compare, verify, finish; ring sizes are "
                          "cached per .locs). 935 M candidate pairs get a 32-symbol set test of 3 LOP3 + POPC each: the "
                          "compare kernel is bound by integer issue (logic pipe 84 %, POPC pipe 76 % busy), not by HBM "
                          "(DESIGN.md 4.5)",
            "cpu_port_s_scaled": t_cpu * N / n_c, "cpu_sample": "first %d rows (%d wells) of the tile, C restatement, 1 thread, "
            "scaled by wells" % (rows, n_c), "cropped_tile_matches_oracle": bool(np.array_equal(got, want))}


def config5(eng, args):
    from oracle import c_port as CP
    n = synth.NOVASEQ_WELLS
    row = 1600
    X, Y = synth.hex_lattice(n, row)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    random.seed(13)
    centres = np.array(random.sample(range(n), 2500), dtype=np.uint32)
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    split = 5                     # window = cycles 20..69: the first 5 are written with every well (cbcl_read.py:77-80)
    D = 4
    tiles = []
    for k in range(D):
        td = synth.make_tile_fast(777 + k, n, NCYC, row)
        pf = (td.filt & 1).astype(bool)
        planes, kinds, nb = [], [], []
        for c in range(NCYC):
            nib = synth.bcl_to_nibbles(td.planes[c])
            if c >= split:
                nib = nib[pf]
            planes.append(synth.pack_nibbles(nib))
            kinds.append("cbcl_excl" if c >= split else "cbcl")
            nb.append(nib.size)
        tiles.append((planes, kinds, nb, td.filt))
    T = args.cbcl_tiles
    t0 = time.perf_counter()
    for s in range(T):
        planes, kinds, nb, filt = tiles[s % D]
        eng.tile_begin(s, n, NCYC)
        eng.tile_put_filter(s, filt)
        for c in range(NCYC):
            eng.tile_put_cbcl(s, c, planes[c], nb[c], kinds[c] == "cbcl_excl")
    eng.sync()
    t_stage = time.perf_counter() - t0
    order = list(range(NCYC))
    _, cnt = eng.count(0, T, order, 2, False, mode=0, per_target=False)      # also builds the PF rank tables (K3)
    t = sync_time(eng, lambda: eng.count_async(0, T, order, 2, False, mode=0), 5)
    planes, kinds, nb, filt = tiles[0]
    t0 = time.perf_counter()
    _, want = CP.count_tile(planes, kinds, filt, centres, offs, idx, 5, 2, False, want_per_target=False)
    t_cpu = time.perf_counter() - t0
    plane_bytes = sum(p.size for p in tiles[0][0])
    return {"config": "5: NovaSeq-style CBCL, %d tiles x %d wells resident in HBM (a lane has 704), 2500 targets x 5 rings, "
                      "cycles 20-69 (5 with every well, 45 with pass-filter wells only)" % (T, n),
            "count_ms": 1e3 * t, "targets_per_s": T * 2500 / t, "wells_compared_per_s": int(cnt[:, 1::5].sum()) / t,
            "tiles": T, "lane_of_704_tiles_ms_scaled": 1e3 * t * 704 / T,
            "staging_s": t_stage, "staged_bytes": int(T * (plane_bytes + n)),
            "cpu_port_tile_s": t_cpu, "cpu_port_targets_per_s_1_thread": 2500 / t_cpu,
            "first_tile_matches_oracle": bool(np.array_equal(cnt[0], want))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="4,3,5")
    ap.add_argument("--cbcl-tiles", type=int, default=352)
    args = ap.parse_args()
    eng = Engine(0)
    fns = {"3": config3, "4": config4, "5": config5}
    for c in args.configs.split(","):
        t0 = time.perf_counter()
        line = fns[c](eng, args)
        line["measure_wall_s"] = time.perf_counter() - t0
        print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
