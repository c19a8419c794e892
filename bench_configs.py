"""bench.py --config exhaustive | stage1 | cbcl: BASELINE.json configs 3, 4 and 5.

Same JSON shape as the headline line of bench.py (value with inputs resident in
HBM, e2e from host buffers through the C ABI, roofline for the dominant kernel,
cpu_baseline = the C restatement of the reference on the box's host cores, and a
parity flag against that oracle).  One GPU: these configs are single-tile
(exhaustive, stage 1) or one lane (CBCL) workloads; a flowcell shards them like
the headline config.

  exhaustive  config 3: every well of one HiSeq-shaped tile is a target out to ring 5
  stage1      config 4: .locs of a full tile -> rings of 2500 sampled targets
  cbcl        config 5: a NovaSeq-style lane of 704 tiles, 4-bit CBCL planes (the first
              compared cycles hold every well, the later ones pass-filter wells only)
"""
import os
import random
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

N = 4309650
ROW = 1571
NCYC = 50
LEVELS = 5
EDIT = 2
SEED = 20261018


def _peak():
    from bench import load_peaks
    return load_peaks()


def parse_sweep(item):
    """'8,4[,16] head_planes=3 ...' -> keyword arguments of Engine.set_tuning."""
    kw = {}
    for part in item.split():
        if "=" in part:
            k, v = part.split("=", 1)
            kw[k] = int(v)
        else:
            f = [int(x) for x in part.split(",")]
            kw["step0"], kw["step1"] = f[0], f[1]
            if len(f) > 2:
                kw["centre_chunk"] = f[2]
    return kw


def _event_time(torch, stream, fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


def _base(args, metric, value, unit, seconds, workload, **cfg):
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": 1e3 * seconds, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64 bit-planes", "data": "synthetic", "config": dict(workload=workload, seed=SEED, **cfg)}


def _engine():
    import torch
    from well_duplicates_b200.engine import Engine
    torch.cuda.set_device(0)
    eng = Engine(0)
    stream = torch.cuda.Stream(device=0)
    eng.set_stream(stream.cuda_stream)
    return torch, eng, stream


# ---------------------------------------------------------------------------------- config 3 --
def _exhaustive_inputs():
    from well_duplicates_b200 import synth
    X, Y = synth.hex_lattice(N, ROW)
    td = synth.make_tile_fast(SEED, N, NCYC, ROW)
    return X, Y, td


def _exhaustive_cpu(X, Y, td, rows, threads, hamming):
    """The C port on `threads` crops of `rows` lattice rows each (every well of a crop a target)."""
    from oracle import c_port as CP
    n_c = rows * ROW
    order = list(range(NCYC))

    def one(k):
        lo = k * n_c
        return CP.count_exhaustive(X[lo:lo + n_c], Y[lo:lo + n_c], [td.planes[c][lo:lo + n_c] for c in order], ["bcl"] * NCYC,
                                   td.filt[lo:lo + n_c], LEVELS, EDIT, hamming)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        res = list(pool.map(one, range(threads)))
    return time.perf_counter() - t0, n_c, res


def run_exhaustive(args):
    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import PinnedArray
    torch, eng, stream = _engine()
    X, Y, td = _exhaustive_inputs()
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    pin = PinnedArray((NCYC, N))
    pin.array[:] = td.planes
    fpin = PinnedArray((N,))
    fpin.array[:] = td.filt
    order = list(range(NCYC))

    def put():
        eng.tile_begin(0, N, NCYC)
        eng.tile_put_filter(0, fpin.array)
        for c in range(NCYC):
            eng.tile_put_bcl(0, c, pin.array[c])
    put()
    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            cnt = eng.count_exhaustive(0, order, LEVELS, EDIT, args.hamming)
        l0 = eng.launch_count()
        t = _event_time(torch, stream, lambda: eng.count_exhaustive(0, order, LEVELS, EDIT, args.hamming), args.steps)
        launches = eng.launch_count() - l0

        def e2e_step():
            put()
            eng.count_exhaustive(0, order, LEVELS, EDIT, args.hamming)
        e2e_step()
        t_e2e = _event_time(torch, stream, e2e_step, max(1, args.e2e_steps))
    # parity + CPU baseline: crops of the same tile, every well of a crop a target, on all host threads
    cores = os.cpu_count() or 1
    rows = 40
    dt, n_c, res = _exhaustive_cpu(X, Y, td, rows, cores, args.hamming)
    eng.load_locs(synth.xy_to_locs_floats(X[:n_c], Y[:n_c]))
    eng.tile_begin(1, n_c, NCYC)
    eng.tile_put_filter(1, td.filt[:n_c])
    for c in range(NCYC):
        eng.tile_put_bcl(1, c, td.planes[c][:n_c])
    got = eng.count_exhaustive(1, order, LEVELS, EDIT, args.hamming)
    ok = bool(np.array_equal(got, res[0]))
    peak, kind = _peak()
    alg = N * NCYC + N * 8 + N + 2 * N * 24                     # planes, locs, filter, packed write + read
    line = _base(args, "targets/sec, exhaustive mode: every well of a HiSeq-shaped tile a target out to ring 5", N / t,
                 "targets/s", t, "exhaustive_tile: %d wells, every well a target x %d rings, %d-cycle BCL substring, %s e=%d" % (
                     N, LEVELS, NCYC, "Hamming" if args.hamming else "Levenshtein", EDIT),
                 l2="one tile (215 MB of planes) per step: larger than L2")
    line.update({
        "wells_compared_per_s": int(cnt[1::5].sum()) / t, "valid_targets": int(cnt[0]), "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak,
                     "traffic": None, "peak_source": "MEASURED_PEAKS.json (%s)" % kind, "kernel": "exh_compare_kernel (+ dense pack, prefix, verify, finish)",
                     "algorithmic_bytes_per_launch": alg,
                     "note": "whole wd_count_exhaustive call. HBM is not what limits it: 935 M candidate pairs per tile get a "
                             "32-symbol set test (3 LOP3 + POPC); the compare kernel runs the integer pipes at 76-84 % "
                             "(profiles/r01_exhaustive_v3_ncu.txt, DESIGN.md 4.5)"},
        "e2e": {"value": N / t_e2e, "unit": "targets/s", "ms_per_step": 1e3 * t_e2e, "h2d_bytes_per_step": int(N * (NCYC + 1)),
                "d2h_bytes_per_step": (1 + 5 * LEVELS) * 8, "staging": "wd_tile_put_bcl of every plane from pinned host memory, then wd_count_exhaustive"},
        "cpu_baseline": {"value": cores * n_c / dt, "unit": "targets/s", "cores": cores, "kind": "port", "seconds": dt,
                         "sample": "%d crops of %d lattice rows (%d wells each, every well a target), one per host thread; C "
                                   "restatement of prepare_cluster_indexes + count_well_duplicates (oracle/welldup_oracle.c)" % (cores, rows, n_c)},
        "counters_match_oracle": ok, "counters_match_note": "the first crop counted by wd_count_exhaustive equals the C oracle's counters"})
    return line


# ---------------------------------------------------------------------------------- config 4 --
def run_stage1(args):
    from oracle import c_port as CP
    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import PinnedArray
    torch, eng, stream = _engine()
    X, Y = synth.hex_lattice(N, ROW)
    xy = synth.xy_to_locs_floats(X, Y)
    pin = PinnedArray((N, 2), np.float32)
    pin.array[:] = xy
    random.seed(13)
    centres = np.array(random.sample(range(N), 2500), dtype=np.uint32)
    with torch.cuda.stream(stream):
        eng.load_locs(pin.array)
        offs, idx = eng.ring_query(centres, LEVELS)
        for _ in range(max(3, args.warmup)):
            eng.ring_query(centres, LEVELS)
        l0 = eng.launch_count()
        t_q = _event_time(torch, stream, lambda: eng.ring_query(centres, LEVELS), args.steps)
        launches = eng.launch_count() - l0

        def e2e_step():
            eng.load_locs(pin.array)
            eng.ring_query(centres, LEVELS)
        t_e2e = _event_time(torch, stream, e2e_step, max(2, args.e2e_steps))
    cores = os.cpu_count() or 1
    per = 8
    chunks = [centres[k * per:(k + 1) * per] for k in range(cores)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        res = list(pool.map(lambda c: CP.rings_csr(X, Y, c), chunks))
    dt = time.perf_counter() - t0
    woffs = np.concatenate([[0]] + [r[0][1:].astype(np.int64) + sum(int(q[0][-1]) for q in res[:i]) for i, r in enumerate(res)])
    widx = np.concatenate([r[1] for r in res])
    ok = bool(np.array_equal(offs[:woffs.size], woffs) and np.array_equal(idx[:widx.size], widx))
    peak, kind = _peak()
    alg_q = 2500 * 410 * 16 + int(idx.size) * 4                  # ~410 grid records of 16 B per target, CSR out
    alg_e = N * 8 + N * 8 + N * 16 + int(idx.size) * 4           # locs in, pixels out, cell key + record, CSR out
    line = _base(args, "targets/sec, prepare_cluster_indexes neighbourhood build on a full-tile .locs", 2500 / t_q, "targets/s", t_q,
                 "stage1: %d-well .locs, 2500 sampled targets x %d rings (grid resident; e2e rebuilds it from host floats)" % (N, LEVELS),
                 l2="value: the grid (69 MB) fits L2 -- the query is latency-bound; e2e streams the 34.5 MB .locs")
    line.update({
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": alg_q / t_q / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_q / t_q / 1e9 / peak,
                     "traffic": None, "peak_source": "MEASURED_PEAKS.json (%s)" % kind, "kernel": "ring_query_kernel<count> + scan + <fill>",
                     "algorithmic_bytes_per_launch": alg_q,
                     "note": "2500 targets are 18 MB of grid records: launch- and latency-bound (two kernels, a scan and a D2H "
                             "of the offsets between them), far from the HBM roofline by construction; the reference's loop "
                             "for the same list takes 181 s"},
        "e2e": {"value": 2500 / t_e2e, "unit": "targets/s", "ms_per_step": 1e3 * t_e2e, "h2d_bytes_per_step": int(N * 8),
                "d2h_bytes_per_step": int(idx.size * 4 + offs.size * 4), "algorithmic_bytes": alg_e, "achieved_gb_per_s": alg_e / t_e2e / 1e9,
                "staging": "wd_locs_load from pinned floats (H2D + K0 pixels + K1 grid build) + wd_ring_query, CSR back on the host"},
        "cpu_baseline": {"value": cores * per / dt, "unit": "targets/s", "cores": cores, "kind": "port", "seconds": dt,
                         "sample": "%d of the 2500 centres per host thread, C restatement of get_indexes (the +-20000-record scan)" % per},
        "counters_match_oracle": ok, "counters_match_note": "level offsets and well lists of the first %d targets equal the C oracle's" % (cores * per)})
    return line


# ---------------------------------------------------------------------------------- config 5 --
def _cbcl_tiles(D, n, row):
    from well_duplicates_b200 import synth
    split = 5                     # window = cycles 20..69: the first 5 are written with every well (cbcl_read.py:77-80)
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
    return tiles


def run_cbcl(args):
    from oracle import c_port as CP
    from well_duplicates_b200 import _lib, synth
    from well_duplicates_b200.engine import PinnedArray
    torch, eng, stream = _engine()
    n, row = synth.NOVASEQ_WELLS, 1600
    X, Y = synth.hex_lattice(n, row)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    random.seed(13)
    centres = np.array(random.sample(range(n), 2500), dtype=np.uint32)
    offs, idx = eng.ring_query(centres, LEVELS)
    eng.load_targets(centres, offs, idx, LEVELS)
    D, T = 4, args.cbcl_tiles
    tiles = _cbcl_tiles(D, n, row)
    kind_code = {"cbcl": _lib.PLANE_CBCL, "cbcl_excl": _lib.PLANE_CBCL_EXCL}

    def put(T):
        for s in range(T):
            planes, kinds, nb, filt = tiles[s % D]
            eng.tile_begin(s, n, NCYC)
            eng.tile_put_filter(s, filt)
            for c in range(NCYC):
                eng.tile_put_cbcl(s, c, planes[c], nb[c], kinds[c] == "cbcl_excl")
    t0 = time.perf_counter()
    put(T)
    eng.sync()
    t_stage = time.perf_counter() - t0
    order = list(range(NCYC))
    with torch.cuda.stream(stream):
        _, cnt = eng.count(0, T, order, EDIT, args.hamming, mode=0, per_target=False)      # also builds the PF rank tables (K3)
        for _ in range(max(3, args.warmup)):
            eng.count_async(0, T, order, EDIT, args.hamming, mode=0)
        l0 = eng.launch_count()
        t = _event_time(torch, stream, lambda: eng.count_async(0, T, order, EDIT, args.hamming, mode=0), args.steps)
        launches = eng.launch_count() - l0
        # wd_set_tuning sweeps in this process (--sweep-steps, as in the lane config)
        sweep, schedules = {}, [x for x in getattr(args, "sweep_steps", "").split(";") if x.strip()]
        for sch in schedules:
            eng.set_tuning(**parse_sweep(sch))
            eng.count_async(0, T, order, EDIT, args.hamming, mode=0)
            sweep[sch] = {"resident_ms": 1e3 * _event_time(torch, stream, lambda: eng.count_async(0, T, order, EDIT, args.hamming, mode=0), args.steps)}
        eng.set_tuning()
        sectors, lines = eng.trace_sectors(0, D, order, EDIT, args.hamming)
        # e2e: inflated blocks in page-locked host memory, laid out [tile][plane][stride] as the staging pipeline leaves
        # them (staging.py), mapped (wd_tile_map_host) and counted; Z distinct host tiles so that no tile slot is
        # served from lines another slot brought into L2
        stride = ((n + 1) // 2 + 255) // 256 * 256
        Z, TPB = min(T, 64), 16
        blocks = [PinnedArray((min(TPB, Z - b0), NCYC, stride)) for b0 in range(0, Z, TPB)]
        fpins = [PinnedArray((n,)) for _ in range(Z)]
        for z in range(Z):
            planes, kinds, nb, filt = tiles[z % D]
            for c in range(NCYC):
                blocks[z // TPB].array[z % TPB, c, :planes[c].size] = planes[c]
            fpins[z].array[:] = filt
        kinds_arr = [np.array([kind_code[k] for k in tiles[d][1]], np.uint8) for d in range(D)]
        nb_arr = [np.array(tiles[d][2], np.uint32) for d in range(D)]

        def e2e_step():
            for s in range(T):
                z = s % Z
                eng.tile_map_host(s, n, blocks[z // TPB].array[z % TPB], kinds=kinds_arr[z % D], n_block=nb_arr[z % D],
                                  pinned_filter=fpins[z].array)
            return eng.count(0, T, order, EDIT, args.hamming, mode=0, per_target=False)[1]
        ok_e2e = bool(np.array_equal(e2e_step(), cnt)) if (T % Z == 0 or Z % D == 0) else None
        t_e2e = _event_time(torch, stream, e2e_step, max(1, args.e2e_steps))
        dma_bytes = eng.last_count_h2d_bytes()
        for sch in schedules:
            eng.set_tuning(**parse_sweep(sch))
            e2e_step()
            sweep[sch]["zero_copy_ms"] = 1e3 * _event_time(torch, stream, e2e_step, max(1, args.e2e_steps))
        eng.set_tuning()
    reps = np.bincount(np.arange(T) % D, minlength=D)
    plane_sectors = int((sectors.sum(axis=1) * reps).sum())
    need = plane_sectors * 32 + int(np.unique(centres >> 5).size) * 32 * T + int(idx.size + centres.size) * 5 + T * (1 + 5 * LEVELS) * 8
    # every distinct tile against the oracle (each of the T slots holds one of them), all host threads as the CPU baseline
    cores = os.cpu_count() or 1

    def one(k):
        planes, kinds, nb, filt = tiles[k % D]
        return CP.count_tile(planes, kinds, filt, centres, offs, idx, LEVELS, EDIT, args.hamming, want_per_target=False)[1]
    n_cpu = 4 * cores
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        res = list(pool.map(one, range(n_cpu)))
    dt = time.perf_counter() - t0
    ok = all(np.array_equal(cnt[s], res[s % D]) for s in range(T))
    peak, kind = _peak()
    plane_bytes = sum(p.size for p in tiles[0][0])
    line = _base(args, "targets/sec, NovaSeq-style CBCL lane (4-bit base+quality bins), full-lane duplicate count", T * 2500 / t, "targets/s", t,
                 "cbcl_lane: %d tiles x %d wells, 2500 targets x %d rings, cycles 20-69 (5 planes with every well, 45 with "
                 "pass-filter wells only), %s e=%d" % (T, n, LEVELS, "Hamming" if args.hamming else "Levenshtein", EDIT),
                 tiles=T, distinct_tiles=D, l2="%.1f GB of planes resident: far larger than L2" % (T * plane_bytes / 1e9))
    line.update({
        "wells_compared_per_s": int(cnt[:, 1::5].sum()) / t, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": need / t / 1e9, "peak": peak, "unit": "GB/s", "frac": need / t / 1e9 / peak,
                     "traffic": None, "peak_source": "MEASURED_PEAKS.json (%s)" % kind, "kernel": "fused_count_kernel<ALL_BCL=false>",
                     "needed_bytes_per_launch": need, "needed_plane_sectors": plane_sectors,
                     "needed_plane_lines_128B": int((lines.sum(axis=1) * reps).sum()),
                     "needed_note": "measured in this run by wd_count_trace_sectors (distinct 32-byte sectors the kernel asks "
                                    "for, per tile and compared position) + filter, index and counter bytes"},
        "e2e": {"value": T * 2500 / t_e2e, "unit": "targets/s", "ms_per_step": 1e3 * t_e2e,
                "h2d_bytes_per_step": int(dma_bytes), "d2h_bytes_per_step": int(T * (1 + 5 * LEVELS) * 8),
                "host_bytes_mapped_per_step": int(T * (plane_bytes + n)), "distinct_host_tiles": Z, "counters_equal_resident_run": ok_e2e,
                "staging": "wd_tile_map_host: the inflated blocks stay in page-locked host memory; wd_count copies the planes of "
                           "the first 2 compared cycles by DMA (h2d_bytes_per_step) and pulls the sectors it needs of the others "
                           "across PCIe; K3 (PF rank of every tile) runs inside the step"},
        "staged_copy_s": t_stage, "staged_copy_note": "wd_tile_put_cbcl of every block from pageable memory (%.1f GB), once, outside value" % (T * (plane_bytes + n) / 1e9),
        "cpu_baseline": {"value": n_cpu * 2500 / dt, "unit": "targets/s", "cores": cores, "kind": "port", "seconds": dt,
                         "sample": "%d tiles (one per host thread at a time) of the same lane, blocks already inflated in RAM; C "
                                   "restatement of the reference incl. its filter-offset table (oracle/welldup_oracle.c)" % n_cpu},
        "sweep_steps": sweep or None,
        "counters_match_oracle": bool(ok), "counters_match_note": "all %d tile rows equal the C oracle's rows of the distinct tile they hold" % T})
    return line


def run(args):
    return {"exhaustive": run_exhaustive, "stage1": run_stage1, "cbcl": run_cbcl}[args.config](args)


def run_reference(args):
    """--impl reference for these configs: the C port on all host threads, bounded sample per step."""
    from oracle import c_port as CP
    from well_duplicates_b200 import synth
    cores = os.cpu_count() or 1
    total_t, units = 0.0, 0
    if args.config == "exhaustive":
        X, Y, td = _exhaustive_inputs()
        for i in range(args.warmup + args.steps):
            dt, n_c, _ = _exhaustive_cpu(X, Y, td, 20, cores, args.hamming)
            if i >= args.warmup:
                total_t += dt
                units += cores * n_c
        sample = "%d crops of 20 lattice rows per step, every well a target, one per host thread" % cores
        metric = "targets/sec, exhaustive mode: every well of a HiSeq-shaped tile a target out to ring 5"
    elif args.config == "stage1":
        X, Y = synth.hex_lattice(N, ROW)
        random.seed(13)
        centres = np.array(random.sample(range(N), 2500), dtype=np.uint32)
        per = 8
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=cores) as pool:
                list(pool.map(lambda k: CP.rings_csr(X, Y, centres[k * per:(k + 1) * per]), range(cores)))
            if i >= args.warmup:
                total_t += time.perf_counter() - t0
                units += cores * per
        sample = "%d centres per host thread per step, C restatement of get_indexes" % per
        metric = "targets/sec, prepare_cluster_indexes neighbourhood build on a full-tile .locs"
    else:
        n, row = synth.NOVASEQ_WELLS, 1600
        X, Y = synth.hex_lattice(n, row)
        random.seed(13)
        centres = np.array(random.sample(range(n), 2500), dtype=np.uint32)
        offs, idx = CP.rings_csr(X, Y, centres)
        tiles = _cbcl_tiles(2, n, row)

        def one(k):
            planes, kinds, nb, filt = tiles[k % 2]
            return CP.count_tile(planes, kinds, filt, centres, offs, idx, LEVELS, EDIT, args.hamming, want_per_target=False)[1]
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=cores) as pool:
                list(pool.map(one, range(2 * cores)))
            if i >= args.warmup:
                total_t += time.perf_counter() - t0
                units += 2 * cores * 2500
        sample = "%d CBCL tiles per step (blocks inflated in RAM), one per host thread at a time" % (2 * cores)
        metric = "targets/sec, NovaSeq-style CBCL lane (4-bit base+quality bins), full-lane duplicate count"
    value = units / total_t
    return {"impl": "reference", "metric": metric, "value": value, "unit": "targets/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64 bit-planes", "data": "synthetic", "config": {"workload": args.config},
            "cpu_baseline": {"value": value, "unit": "targets/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "targets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
