#!/usr/bin/env python3
"""Benchmark of the well_duplicates hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config lane|exhaustive|stage1|cbcl] [--from-files]

Default workload (``--config lane``, BASELINE.json configs[1] split per GPU; weak
scaling, N GPUs = N lanes = the 8 x 96-tile flowcell at N = 8): one synthetic
HiSeq-4000 lane of 96 tiles x 4 309 650 wells, 2500 sampled targets out to ring 5,
a 50-cycle substring from BCL byte planes, default (Levenshtein, e = 2) compare.
A step is one pass of the hot path (fused gather-decode + compare + counter
reduction, then K7 + the lane-counter all-reduce when N > 1) over all of the
rank's tiles.

* ``value``: planes resident in HBM, CUDA events on the launching stream, max over ranks.
* ``e2e``: every step starts from planes in pinned HOST memory and ends with the
  counters on the host, through the C ABI (wd_tile_map_host + wd_count: head planes
  by DMA, the rest pulled as sectors across PCIe); ``e2e_logged`` is the same step
  with the duplicate-pair log the production flag set asks for (no -q); ``e2e_staged``
  copies every plane to HBM first.  ``--from-files`` adds ``e2e_files``: .filter/.bcl.gz
  files on local disk -> counters through staging.lane_batches.  Gunzip is outside
  value and e2e (BASELINE.json north_star) and reported beside them.
* ``roofline``: ``achieved`` = bytes the algorithm NEEDS per launch / event time.  The
  numerator is measured in this run by wd_count_trace_sectors: the fused kernel in a
  build that records its reads, i.e. the distinct 32-byte sectors per plane it asks for
  with its early exits (+ filter, index and counter bytes).  ``traffic`` (ncu DRAM bytes)
  is reported only when profiles/traffic.json was captured from the very kernel sources
  that are running (hash check); otherwise null.
* ``--impl reference``: the C restatement of the reference (oracle/) on all host
  threads, on a bounded sample of the same workload.

``--config exhaustive|stage1|cbcl`` time BASELINE.json configs 3, 4, 5 (bench_configs.py)
and print the same JSON shape (value, roofline, cpu_baseline, parity flag).
"""
import argparse
import hashlib
import json
import os
import random
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from bench_configs import parse_sweep

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WELLS = 4309650
STRIDE = (N_WELLS + 255) // 256 * 256
ROW_LEN = 1571
N_CYCLES = 50
N_TARGETS = 2500
LEVELS = 5
EDIT = 2
TILES_PER_LANE = 96
SEED = 20261018
METRIC = "targets/sec (with wells compared/sec) on synthetic HiSeq 4000 lanes, 96 tiles x 2500 targets per GPU"
# profiles/r02_fetch_granularity_micro.txt: scattered loads that miss L2 complete at 46 G/s on this part whatever
# the fill size (64 or 128 bytes) -- the request-rate ceiling the gather runs against
RANDOM_LINE_REQUESTS_PER_S = 46.0e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="lane", choices=["lane", "exhaustive", "stage1", "cbcl"],
                    help="lane: BASELINE configs[1] per GPU (the headline); exhaustive / stage1 / cbcl: configs 3 / 4 / 5")
    ap.add_argument("--tiles", type=int, default=TILES_PER_LANE, help="tiles per GPU")
    ap.add_argument("--distinct-tiles", type=int, default=12,
                    help="distinct synthetic tiles kept in pinned host memory (reused round-robin for the tile slots)")
    ap.add_argument("--mode", type=int, default=0, help="0 fused kernel, 1 two-pass kernels, 2 fused + duplicate-pair log")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-mode", default="both", choices=["staged", "zerocopy", "both"],
                    help="staged: every plane copied to HBM through wd_tile_put_bcl; zerocopy: planes stay in pinned "
                         "host memory (wd_tile_map_host) and the kernel pulls the sectors it needs over PCIe")
    ap.add_argument("--cpu-tiles", type=int, default=0, help="tiles in the cpu_baseline sample (0 = 24 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inflate", action="store_true", help="skip the host gunzip measurement (reported beside the metric)")
    ap.add_argument("--hamming", action="store_true")
    ap.add_argument("--zc-blocks", type=int, default=0,
                    help="distinct pinned host tiles for the zero-copy e2e (0 = one per tile slot if the host can pin them)")
    ap.add_argument("--sweep-steps", default="", help="wd_set_tuning sweeps timed in this process, ';'-separated: "
                                                      "'8,4' = rounds of 8 then 4 cycles, optional ',centre_chunk' and "
                                                      "'name=value' items (head_planes, head_groups, visit_order)")
    ap.add_argument("--no-files", action="store_true", help="skip e2e_files (.filter/.bcl.gz files on local disk -> counters)")
    ap.add_argument("--files-tiles", type=int, default=0, help="tiles per GPU for e2e_files (0 = --tiles: the whole lane)")
    ap.add_argument("--files-dir", default="/tmp/wd_bench_run")
    ap.add_argument("--cbcl-tiles", type=int, default=704, help="--config cbcl: tiles resident (a NovaSeq lane has 704)")
    ap.add_argument("--library", default="", help="A/B measurement: another build of libwelldup.so (python -m well_duplicates_b200.build --tag=...)")
    return ap.parse_args()


def make_targets():
    """Seeded sample + rings on the synthetic lattice (same for every tile, as in
    production, Snakefile.count_dups:168)."""
    from well_duplicates_b200 import synth
    X, Y = synth.hex_lattice(N_WELLS, ROW_LEN)
    random.seed(13)
    centres = np.array(random.sample(range(N_WELLS), N_TARGETS), dtype=np.uint32)
    return X, Y, centres


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self, wait_s=5.0):
        """Starts nvidia-smi -lms 50 and waits for its first line (it can take a second to come up)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
            return
        t0 = time.time()
        while not self.lines and time.time() - t0 < wait_s:
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, windows):
        """windows: [(wall t0, wall t1)] of the timed regions; only samples taken inside one of them count."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6 or not any(t0 <= ts <= t1 + 0.03 for t0, t1 in windows):
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": "samples every 50 ms inside the timed regions (value loop and e2e loops, %.2f s in all)"
                                               % sum(t1 - t0 for t0, t1 in windows)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def kernel_source_hash():
    """sha256 over the CUDA sources of libwelldup.so: ties an ncu capture in profiles/ to the kernels that run."""
    csrc = os.path.join(ROOT, "well_duplicates_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):            # the .cc files are host-only (inflate, NCCL glue, page-locking)
            with open(os.path.join(csrc, name), "rb") as fh:
                h.update(name.encode() + b"\0" + fh.read())
    return h.hexdigest()


def load_traffic(key):
    """Per-launch DRAM bytes from the ncu capture of THESE kernel sources (profiles/traffic.json, written by
    profiles/make_traffic.py together with the source hash); None when the capture is of other sources."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    with open(p) as fh:
        t = json.load(fh)
    if t.get("kernel_source_sha256") != kernel_source_hash():
        return None, "profiles/traffic.json was captured from other kernel sources: ignored"
    return t.get(key), "ncu dram__bytes_read+write of this kernel, same kernel sources (profiles/traffic.json)"


def needed_bytes(eng, D, n_tiles, order, centres, offs, idx, hamming):
    """What one launch over n_tiles tiles has to bring in from HBM, measured: wd_count_trace_sectors runs the fused
    kernel (same schedule, same early exits) in a build that records every plane read, on the D distinct resident
    tiles; tile slot s holds distinct tile s % D."""
    sectors, lines = eng.trace_sectors(0, D, order, EDIT, hamming)          # [D, cycles]
    reps = np.bincount(np.arange(n_tiles) % D, minlength=D)
    plane_sectors = int((sectors.sum(axis=1) * reps).sum())
    plane_lines = int((lines.sum(axis=1) * reps).sum())
    filt_sectors = int(np.unique(centres >> 5).size) * n_tiles             # every target's centre PF byte
    # index arrays are the same for every tile of the launch (they stay in L2 after the first): once per launch
    n_slots = int(idx.size + centres.size)
    index_bytes = n_slots * 5 + centres.size * (8 + 4 + 4 * LEVELS)        # slot_well + slot_level; tgt_off, visit, level_len
    out_bytes = n_tiles * (1 + 5 * LEVELS) * 8
    total = plane_sectors * 32 + filt_sectors * 32 + index_bytes + out_bytes
    return {"bytes": int(total), "plane_sectors": plane_sectors, "plane_lines_128B": plane_lines,
            "sectors_per_position": [int(x) for x in (sectors * reps[:, None]).sum(axis=0)],
            "filter_sectors": filt_sectors, "index_bytes": int(index_bytes), "counter_bytes": int(out_bytes)}


def cpu_baseline_sample(planes_by_tile, filts, centres, offs, idx, n_tiles, threads, hamming, tile_ids=None):
    """C port of the reference's per-tile path (oracle/welldup_oracle.c), one tile per thread."""
    from oracle import c_port as CP
    CP.lib()

    def one(k):
        td = planes_by_tile[k % len(planes_by_tile)]
        _, c = CP.count_tile([td[c] for c in range(N_CYCLES)], ["bcl"] * N_CYCLES, filts[k % len(filts)], centres,
                             offs, idx, LEVELS, EDIT, hamming, want_per_target=False)
        return c
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        res = list(pool.map(one, tile_ids if tile_ids is not None else range(n_tiles)))
    dt = time.perf_counter() - t0
    wells = int(sum(int(r[1::5].sum()) for r in res))
    return dt, wells, res


def gz_member(raw):
    """One gzip member (level 1) holding raw."""
    import struct
    import zlib
    comp = zlib.compress(raw, 1)
    return (b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + comp[2:-4] +
            struct.pack("<II", zlib.crc32(raw) & 0xffffffff, len(raw) & 0xffffffff))


def host_inflate_sample(plane, threads, seconds=4.0):
    """Gunzip stays on the host (BASELINE.json north_star) and is reported beside the metric: the product's
    inflate (wd_inflate_batch, csrc/wd_inflate.cc -- what staging.Stager runs) on .bcl.gz-shaped members
    (one compressed plane of the benchmark tile) on all host threads, and zlib -- what the reference's
    gzip.open().read() runs -- on the same members."""
    import struct
    import zlib
    from well_duplicates_b200 import _lib
    raw = struct.pack("<I", plane.size) + plane.tobytes()
    comp = gz_member(raw)
    lib = _lib.load()
    src = np.frombuffer(comp, np.uint8)
    outs = np.empty((threads, len(raw)), np.uint8)
    jobs = (_lib.InflateJob * threads)()
    for k in range(threads):
        jobs[k].src, jobs[k].size, jobs[k].dst, jobs[k].dst_cap = src.ctypes.data, len(comp), outs[k].ctypes.data, len(raw)
    assert lib.wd_inflate_batch(jobs, threads, threads) == 0 and outs[threads - 1].tobytes() == raw
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < seconds / 2:
        assert lib.wd_inflate_batch(jobs, threads, threads) == 0
        done += threads * len(raw)
    dt = time.perf_counter() - t0

    def one(_):
        return len(zlib.decompressobj(wbits=31).decompress(comp))
    z0 = time.perf_counter()
    zdone = 0
    with ThreadPoolExecutor(max_workers=threads) as pool:
        while time.perf_counter() - z0 < seconds / 2:
            zdone += sum(pool.map(one, range(threads)))
    zdt = time.perf_counter() - z0
    return {"gb_per_s": done / dt / 1e9, "threads": threads, "seconds": dt, "compression_ratio": len(comp) / len(raw),
            "lane_seconds": TILES_PER_LANE * N_CYCLES * (N_WELLS + 4) / (done / dt),
            "zlib_gb_per_s": zdone / zdt / 1e9, "zlib_lane_seconds": TILES_PER_LANE * N_CYCLES * (N_WELLS + 4) / (zdone / zdt),
            "note": "wd_inflate_batch (the product's staging inflate, CRC-32 checked) on gzip members shaped like one "
                    ".bcl.gz plane of the benchmark tile (level 1), one member per host thread at a time; zlib_* = the "
                    "same members through zlib, which the reference's gzip.open().read() uses; lane_seconds = the 96 x "
                    "50 planes of one lane at that rate. Outside value and e2e, which start from inflated planes in "
                    "pinned host memory"}


def run_reference(args, rank, world):
    """The reference's CPU path (C port of it) on the box's host cores."""
    if rank != 0:
        return
    if args.config != "lane":
        import bench_configs
        print(json.dumps(bench_configs.run_reference(args)))
        return
    from oracle import c_port as CP
    from well_duplicates_b200 import synth
    cores = os.cpu_count() or 1
    X, Y, centres = make_targets()
    offs, idx = CP.rings_csr(X, Y, centres)
    n_distinct = min(cores, 8)
    tiles = [synth.make_tile_fast(SEED + k, N_WELLS, N_CYCLES, ROW_LEN) for k in range(n_distinct)]
    planes = [t.planes for t in tiles]
    filts = [t.filt for t in tiles]
    per_step = 8 * cores
    for _ in range(args.warmup):
        cpu_baseline_sample(planes, filts, centres, offs, idx, min(per_step, 4), cores, args.hamming)
    total_t, total_tiles, total_wells = 0.0, 0, 0
    for _ in range(args.steps):
        dt, wells, _ = cpu_baseline_sample(planes, filts, centres, offs, idx, per_step, cores, args.hamming)
        total_t += dt
        total_tiles += per_step
        total_wells += wells
    value = total_tiles * N_TARGETS / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "targets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 bit-planes",
        "data": "synthetic", "wells_compared_per_s": total_wells / total_t,
        "config": workload_config(args, per_gpu_tiles=args.tiles),
        "cpu_baseline": {"value": value, "unit": "targets/s", "cores": cores, "kind": "port",
                         "sample": "%d tiles per step (one per host thread) of the same synthetic lane, planes already "
                                   "gunzipped in RAM; C restatement of the reference (oracle/welldup_oracle.c)" % per_step},
        "e2e": {"value": value, "unit": "targets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _ranges(cpus):
    out, start, prev = [], None, None
    for c in cpus + [None]:
        if start is None:
            start = prev = c
        elif c is not None and c == prev + 1:
            prev = c
        else:
            out.append("%d" % start if start == prev else "%d-%d" % (start, prev))
            start = prev = c
    return ",".join(out)


def workload_config(args, per_gpu_tiles):
    return {"workload": "hiseq4000_lane_per_gpu: %d tiles x %d wells, %d targets x %d rings, %d-cycle BCL substring, "
                        "%s e=%d" % (per_gpu_tiles, N_WELLS, N_TARGETS, LEVELS, N_CYCLES,
                                     "Hamming" if args.hamming else "Levenshtein", EDIT),
            "tiles_per_gpu": per_gpu_tiles, "targets_per_tile": N_TARGETS, "levels": LEVELS, "cycles": N_CYCLES,
            "wells_per_tile": N_WELLS, "kernel": {0: "fused", 1: "two-pass", 2: "fused + duplicate-pair log"}[args.mode],
            "distinct_tiles": args.distinct_tiles, "seed": SEED,
            "l2": "inputs (%.1f GB of planes per GPU) are far larger than L2; no flush needed" % (
                per_gpu_tiles * N_WELLS * N_CYCLES / 1e9),
            "parallelism": "tiles sharded ordinal % n_gpus; one int64 all-reduce of the counter rows per step"}


def write_lane_files(run_dir, lane, names, distinct_tiles, threads):
    """A lane of .filter / .bcl.gz files (gzip level 1): D distinct tiles written once, hard-linked under the tile names."""
    import gzip
    from well_duplicates_b200 import synth
    ldir = synth.basecalls_dir(run_dir, lane)
    os.makedirs(ldir, exist_ok=True)
    for c in range(N_CYCLES):
        os.makedirs(os.path.join(ldir, "C%d.1" % (c + 1)), exist_ok=True)
    comp = 0

    def write_plane(job):
        d, c, plane = job
        path = os.path.join(ldir, "C%d.1" % (c + 1), "d%d.src" % d)
        with gzip.open(path, "wb", compresslevel=1) as fh:
            fh.write(synth.bcl_plane_bytes(plane))
        return os.path.getsize(path)

    with ThreadPoolExecutor(max_workers=threads) as pool:
        for d, td in enumerate(distinct_tiles):
            with open(os.path.join(ldir, "d%d.fsrc" % d), "wb") as fh:
                fh.write(synth.filter_file_bytes(td.filt))
            comp += sum(pool.map(write_plane, [(d, c, td.planes[c]) for c in range(N_CYCLES)]))
    for k, name in enumerate(names):
        d = k % len(distinct_tiles)
        links = [("d%d.fsrc" % d, "s_%d_%s.filter" % (lane, name))]
        links += [(os.path.join("C%d.1" % (c + 1), "d%d.src" % d), os.path.join("C%d.1" % (c + 1), "s_%d_%s.bcl.gz" % (lane, name)))
                  for c in range(N_CYCLES)]
        for src, dst in links:
            p = os.path.join(ldir, dst)
            if os.path.lexists(p):
                os.remove(p)
            os.link(os.path.join(ldir, src), p)
    return comp // max(1, len(distinct_tiles))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.library:
        from well_duplicates_b200 import _lib as _binding
        _binding.LIB_PATH = os.path.abspath(args.library)

    # stdout carries ONE JSON line: native libraries write there too (NCCL prints its version banner on
    # fd 1), so keep a private handle on the real stdout and point fd 1 at stderr for the rest of the run
    sys.stdout.flush()
    report = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.config != "lane":
        import bench_configs
        if rank == 0:
            report.write(json.dumps(bench_configs.run(args)) + "\n")
            report.flush()
        return

    import torch
    import torch.distributed as dist

    from well_duplicates_b200 import _lib, synth
    from well_duplicates_b200.engine import Engine, PinnedArray

    torch.cuda.set_device(local)
    if world > 1:
        from well_duplicates_b200.engine import bind_to_gpu_numa_node
        bound = None if os.environ.get("WELLDUP_NO_NUMA_BIND") else bind_to_gpu_numa_node(local)
        sys.stderr.write("rank %d: GPU %d, host cores %s\n" % (rank, local, "unchanged" if bound is None else
                                                               "%s (next to %s)" % (_ranges(bound[1]), bound[0])))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        # the counter all-reduce is the library's own ncclAllReduce (wd_comm_init / wd_allreduce_i64);
        # torch.distributed only carries the 128-byte id, the barriers and the max over ranks of the timings
        uid = [Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)

    # ---- inputs -------------------------------------------------------------------
    X, Y, centres = make_targets()
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, LEVELS)          # stage 1 on the GPU
    eng.load_targets(centres, offs, idx, LEVELS)
    n_tiles = args.tiles
    D = max(1, min(args.distinct_tiles, n_tiles))
    pins = [PinnedArray((N_CYCLES, N_WELLS)) for _ in range(D)]
    with ThreadPoolExecutor(max_workers=min(D, 6)) as pool:
        tds = list(pool.map(lambda k: synth.make_tile_fast(SEED + 1000 * rank + k, N_WELLS, N_CYCLES, ROW_LEN,
                                                           out=pins[k].array), range(D)))
    filt_pins = [PinnedArray((N_WELLS,)) for _ in range(D)]
    for k in range(D):
        filt_pins[k].array[:] = tds[k].filt

    def push_tiles():
        for s in range(n_tiles):
            k = s % D
            eng.tile_begin(s, N_WELLS, N_CYCLES)
            eng.tile_put_filter(s, filt_pins[k].array)
            pl = pins[k].array
            for c in range(N_CYCLES):
                eng.tile_put_bcl(s, c, pl[c])

    push_tiles()
    eng.sync()
    order = list(range(N_CYCLES))
    h2d_per_step = n_tiles * (N_CYCLES * N_WELLS + N_WELLS)

    # multi-GPU bookkeeping: rows of the all-reduce buffer = every tile of every lane + one per lane
    total_tiles = n_tiles * world
    ordinals = np.arange(rank, total_tiles, world)
    tile_row = ordinals.astype(np.int32)
    lane_row = (total_tiles + ordinals // n_tiles).astype(np.int32)
    n_rows = total_tiles + world
    width = 1 + 5 * LEVELS

    def step(fetch, mode=None):
        eng.count_async(0, n_tiles, order, EDIT, args.hamming, mode=args.mode if mode is None else mode, per_target=False)
        if world > 1:
            eng.publish_counters(tile_row, lane_row, n_rows)     # K7
            eng.allreduce_published()                            # ncclAllReduce(int64, sum) over NVLink, own stream
            if fetch:
                return eng.published_fetch(n_rows * width).reshape(n_rows, width)
            return None
        if fetch:
            return eng.count_fetch()[1]
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        eng.comm_join()
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / n

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step(False)
        counters = step(True)
        barrier()
        # clocks are sampled from here to the end of the last timed region (value, then the e2e loops): the device-
        # timed value loop alone is a few milliseconds, shorter than nvidia-smi's sampling period
        sampler = ClockSampler(local)
        sampler.start()
        barrier()                          # every rank's sampler is up: nobody waits for a late rank inside the timed loop
        l0 = eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            step(False)
        eng.comm_join()                    # the last all-reduce is inside the timed region
        ev1.record(stream)
        barrier()
        windows = [(t_w0, time.time())]
        launches = eng.launch_count() - l0
        ms = ev0.elapsed_time(ev1)

        # ---- what the launch needs to read, measured (rank 0; the other ranks hold the same kind of tiles) ------
        need = needed_bytes(eng, D, n_tiles, order, centres, offs, idx, args.hamming) if (rank == 0 and args.mode != 1) else None

        sweep = {}
        schedules = [x for x in args.sweep_steps.split(";") if x.strip()]
        for sch in schedules:
            eng.set_tuning(**parse_sweep(sch))
            step(False)
            sweep[sch] = {"resident_ms": timed(lambda: step(False), args.steps)}
        eng.set_tuning()

        # ---- parity of what was timed: this rank's first tiles against the oracle, and the all-reduced lane rows ---
        my_rows = counters[tile_row] if world > 1 else counters
        check_tiles = list(range(min(2, n_tiles))) if world > 1 else []
        ok_local = True
        if check_tiles:
            _, _, res = cpu_baseline_sample([p.array for p in pins], [t.filt for t in tds], centres, offs, idx, len(check_tiles),
                                            len(check_tiles), args.hamming, tile_ids=check_tiles)
            ok_local = all(np.array_equal(res[i], my_rows[k]) for i, k in enumerate(check_tiles))
        if world > 1:
            lane_ok = all(np.array_equal(counters[total_tiles + ln], counters[ln * n_tiles:(ln + 1) * n_tiles].sum(axis=0))
                          for ln in range(world))
            flag = torch.tensor([1 if (ok_local and lane_ok) else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            multi_ok = bool(int(flag[0]))

        # ---- e2e, zero-copy flavour: planes stay in pinned host memory ------------------
        zc_ms = zc_log_ms = None
        if args.e2e_steps > 0 and args.e2e_mode in ("zerocopy", "both"):
            # The staging pipeline hands over page-locked blocks laid out [tile][plane][stride] (staging.py); the
            # same here: blocks of TPB tiles, one distinct host tile per tile slot so that no slot can be served
            # from lines another slot brought into L2 -- unless the host cannot pin that much (then blocks are shared)
            TPB = 12
            n_zc = n_tiles if args.zc_blocks <= 0 else max(1, min(n_tiles, args.zc_blocks))
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            need_host = n_zc * (N_CYCLES + 1) * STRIDE * local_world
            try:
                import psutil
                if args.zc_blocks <= 0 and psutil.virtual_memory().available < 1.25 * need_host + (8 << 30):
                    n_zc = min(n_tiles, TPB)
            except ImportError:
                pass
            blocks = [PinnedArray((min(TPB, n_zc - b0), N_CYCLES, STRIDE)) for b0 in range(0, n_zc, TPB)]
            zc_filt = [PinnedArray((N_WELLS,)) for _ in range(n_zc)]

            def fill(s):
                blocks[s // TPB].array[s % TPB, :, :N_WELLS] = pins[s % D].array
                zc_filt[s].array[:] = filt_pins[s % D].array
            with ThreadPoolExecutor(max_workers=8) as pool:
                list(pool.map(fill, range(n_zc)))

            def map_tiles():
                for s in range(n_tiles):
                    z = s % n_zc
                    eng.tile_map_host(s, N_WELLS, blocks[z // TPB].array[z % TPB], pinned_filter=zc_filt[z].array)

            map_tiles()
            zc_counters = step(True)
            zc_ok = bool(np.array_equal(zc_counters, counters))
            barrier()
            ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.time()
            ev4.record(stream)
            for _ in range(args.e2e_steps):
                map_tiles()
                res = step(True)
            ev5.record(stream)
            barrier()
            windows.append((w0, time.time()))
            zc_ms = ev4.elapsed_time(ev5) / args.e2e_steps
            d2h_per_step = int(res.size * 8)
            zc_dma_bytes = eng.last_count_h2d_bytes()
            zc_head, zc_dma_rate = eng.last_count_staging()
            # sectors the kernel pulls across PCIe = what it asks for beyond the head planes, measured the same way
            zc_need = None
            if rank == 0:
                sec, _ = eng.trace_sectors(0, min(D, n_tiles), order, EDIT, args.hamming)
                reps = np.bincount(np.arange(n_tiles) % min(D, n_tiles), minlength=min(D, n_tiles))
                head = zc_head
                zc_need = {"pulled_sectors": int((sec[:, head:].sum(axis=1) * reps).sum()),
                           "note": "distinct 32-byte sectors of the planes behind the %d head planes that the kernel "
                                   "reads from host memory (wd_count_trace_sectors on the mapped tiles)" % head}

            # the production flag set has no -q: the same step with every duplicate pair logged and fetched
            # with both sequences (count_well_duplicates.py:258-262), as count_cli does
            def logged_step():
                map_tiles()
                r = step(True, mode=_lib.MODE_FUSED_LOG)
                rows, codes = eng.dup_pairs(with_seqs=True)
                return r, rows
            r_log, rows_log = logged_step()
            log_ok = bool(np.array_equal(r_log, counters)) and len(rows_log) == int(my_rows[:, 2::5].sum())
            w0 = time.time()
            zc_log_ms = timed(logged_step, args.e2e_steps)
            windows.append((w0, time.time()))
            zc_log_pairs = int(len(rows_log))

            def zc_step():
                map_tiles()
                step(True)
            for sch in schedules:
                eng.set_tuning(**parse_sweep(sch))
                zc_step()
                sweep[sch]["zero_copy_ms"] = timed(zc_step, args.e2e_steps)
            eng.set_tuning()
            eng.sync()
            for z in blocks:
                z.free()

        # ---- e2e: host planes -> C ABI -> counters on the host, every step -------------
        e2e_ms = None
        if args.e2e_steps > 0 and args.e2e_mode in ("staged", "both"):
            push_tiles()
            step(True)
            barrier()
            ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.time()
            ev2.record(stream)
            for _ in range(args.e2e_steps):
                push_tiles()
                res = step(True)
            ev3.record(stream)
            barrier()
            windows.append((w0, time.time()))
            e2e_ms = ev2.elapsed_time(ev3) / args.e2e_steps
            d2h_per_step = int(res.size * 8)

        clocks = sampler.stop(windows)

        # ---- from compressed files on local disk, through the staging pipeline --------------------------------
        files = None
        if not args.no_files:
            from well_duplicates_b200 import staging
            from well_duplicates_b200.reader import BCLReader
            ft = args.files_tiles or n_tiles
            names = ["%d%d%02d" % (s, w, t) for s in (1, 2) for w in (1, 2, 3, 4) for t in range(1, 25)][:ft]
            run_dir = "%s_r%d" % (args.files_dir, rank)
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            threads = max(1, (os.cpu_count() or 1) // local_world)
            t0 = time.perf_counter()
            comp_per_tile = write_lane_files(run_dir, rank + 1, names, tds[:min(D, 8)], threads)
            t_write = time.perf_counter() - t0
            rd = BCLReader(run_dir, engine=eng)
            st = staging.Stager(threads=threads)

            def lane_from_files():
                rows, inflated = [], 0
                for got, batch in staging.lane_batches(st, lambda t: rd.get_tile(rank + 1, t), names, order):
                    plane_of = st.deliver(eng, batch, first_slot=0, zero_copy=True)
                    _, cnt = eng.count(0, len(got), [plane_of[c] for c in order], EDIT, args.hamming, mode=0, per_target=False)
                    rows.append(cnt)
                    inflated += batch.inflated_bytes
                return np.concatenate(rows), inflated
            rows_f, _ = lane_from_files()                  # page-locked blocks allocated, page cache warm
            files_ok = all(np.array_equal(rows_f[k], my_rows[k % min(D, 8)]) for k in range(min(ft, min(D, 8)))) if min(D, 8) <= n_tiles else None
            barrier()
            w0 = time.perf_counter()
            rows_f, inflated = lane_from_files()
            wall = time.perf_counter() - w0
            barrier()
            st.close()
            import shutil
            shutil.rmtree(run_dir, ignore_errors=True)
            files = {"wall_s": wall, "tiles": ft, "inflated_bytes": int(inflated), "threads": threads,
                     "compressed_bytes_per_tile": int(comp_per_tile), "write_files_s": t_write, "ok": files_ok}

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms or 0.0, zc_ms or 0.0, zc_log_ms or 0.0, files["wall_s"] if files else 0.0],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        e2e_ms = float(t[1]) if e2e_ms is not None else None
        zc_ms = float(t[2]) if zc_ms is not None else None
        zc_log_ms = float(t[3]) if zc_log_ms is not None else None
        if files:
            files["wall_s"] = float(t[4])

    ms_per_step = ms / args.steps
    targets_per_step = total_tiles * N_TARGETS
    wells_per_step = int(counters[:total_tiles, 1::5].sum())
    value = targets_per_step / (ms_per_step / 1e3)

    if rank == 0:
        peak, peak_kind = load_peaks()
        line = {
            "metric": METRIC, "value": value, "unit": "targets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64 bit-planes", "data": "synthetic",
            "wells_compared_per_s": wells_per_step / (ms_per_step / 1e3),
            "config": workload_config(args, n_tiles),
            "clocks": clocks, "gpu_launches": int(launches),
        }
        if need is not None:
            achieved = need["bytes"] / (ms_per_step / 1e3) / 1e9
            traffic, traffic_note = load_traffic("fused" if args.mode != 1 else "two_pass")
            if traffic is not None:
                traffic = int(traffic * n_tiles / TILES_PER_LANE)
            line["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json (%s)" % peak_kind,
                "kernel": "fused_count_kernel", "needed_bytes_per_launch": need["bytes"],
                "needed": {k: need[k] for k in ("plane_sectors", "plane_lines_128B", "filter_sectors", "index_bytes", "counter_bytes")},
                "needed_sectors_per_position": need["sectors_per_position"],
                "needed_note": "measured in this run: wd_count_trace_sectors = the fused kernel with every plane read "
                               "recorded (same schedule and early exits as the timed launches): distinct 32-byte sectors "
                               "per tile and compared position x 32 B, + the filter sectors of all centres, the index "
                               "arrays once per launch, the counter rows",
                "over_fetch": None if traffic is None else traffic / need["bytes"],
                "dram_frac": None if traffic is None else traffic / (ms_per_step / 1e3) / 1e9 / peak,
                "request_bound": {
                    "lines_128B_per_launch": need["plane_lines_128B"],
                    "achieved_g_lines_per_s": need["plane_lines_128B"] / (ms_per_step / 1e3) / 1e9,
                    "random_access_ceiling_g_per_s": RANDOM_LINE_REQUESTS_PER_S / 1e9,
                    "note": "every distinct 128-byte line of a plane is one request to DRAM; scattered requests complete "
                            "at 46 G/s on this part whatever their fill size (profiles/r02_fetch_granularity_micro.txt)"},
                "note": "duration = CUDA events around %d launches on the launching stream; at N>1 it also "
                        "covers the publish kernel and the all-reduce" % args.steps}
        if world > 1:
            line["counters_match_oracle"] = multi_ok
            line["counters_match_note"] = ("every rank: its first %d tiles' all-reduced rows equal the C oracle's, and every lane "
                                           "row equals the sum of its tile rows; flags combined with an all-reduce(min)" % len(check_tiles))
        staged = None if e2e_ms is None else {
            "value": targets_per_step / (e2e_ms / 1e3), "unit": "targets/s", "ms_per_step": e2e_ms,
            "staging": "wd_tile_put_bcl: every plane copied from pinned host memory to HBM, then counted",
            "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": d2h_per_step,
            "h2d_gb_per_s_per_gpu": h2d_per_step / (e2e_ms / 1e3) / 1e9}
        if zc_ms is not None:
            pulled = None if zc_need is None else zc_need["pulled_sectors"] * 32
            line["e2e"] = {
                "value": targets_per_step / (zc_ms / 1e3), "unit": "targets/s", "ms_per_step": zc_ms,
                "staging": "wd_tile_map_host: planes and filters stay in pinned host memory (%.1f GB per step and GPU); "
                           "wd_count copies the plane of the first compared cycle (head_planes_by_dma) to HBM by DMA, tile group after "
                           "tile group, while the counting kernel reads the sectors it needs of the later planes "
                           "across PCIe" % (n_tiles * (N_CYCLES + 1) * N_WELLS / 1e9),
                "h2d_bytes_per_step": int(zc_dma_bytes + (pulled or 0)),
                "h2d_dma_bytes_per_step": int(zc_dma_bytes), "head_planes_by_dma": zc_head, "h2d_dma_gb_per_s": zc_dma_rate,
                "h2d_pulled_bytes_per_step": pulled,
                "h2d_bytes_note": "per GPU. DMA bytes are counted by the library; pulled bytes = 32 B x the distinct sectors "
                                  "behind the head planes that the kernel asks for, measured in this run "
                                  "(wd_count_trace_sectors on the mapped tiles)",
                "pull_requests_per_s_per_gpu": None if pulled is None else pulled / 32 / (zc_ms / 1e3),
                "host_bytes_mapped_per_step": int(n_tiles * (N_CYCLES + 1) * N_WELLS),
                "distinct_host_tiles": n_zc,
                "d2h_bytes_per_step": d2h_per_step, "counters_equal_resident_run": zc_ok}
            line["e2e_logged"] = {
                "value": targets_per_step / (zc_log_ms / 1e3), "unit": "targets/s", "ms_per_step": zc_log_ms,
                "vs_e2e": zc_log_ms / zc_ms, "pairs_logged_per_step_per_gpu": zc_log_pairs, "rows_and_counters_ok": log_ok,
                "what": "the e2e step in WD_MODE_FUSED_LOG + wd_dup_pairs_seqs: every duplicate pair with both sequences "
                        "on the host, what count_well_duplicates.py prints without -q (:258-262)"}
            if staged is not None:
                line["e2e_staged"] = staged
        else:
            line["e2e"] = staged
        if files is not None:
            line["e2e_files"] = {
                "value": files["tiles"] * world * N_TARGETS / files["wall_s"], "unit": "targets/s", "seconds": files["wall_s"],
                "tiles_per_gpu": files["tiles"], "flowcell_seconds_at_this_rate": files["wall_s"] * TILES_PER_LANE / files["tiles"],
                "host_inflate_gb_per_s_per_rank": files["inflated_bytes"] / files["wall_s"] / 1e9,
                "inflate_threads_per_rank": files["threads"], "compressed_bytes_per_tile": files["compressed_bytes_per_tile"],
                "counters_match_resident_run": files["ok"],
                "what": "every rank: its lane's .filter / .bcl.gz files (gzip level 1, page cache warm, written outside the "
                        "timed region) -> staging.lane_batches (native inflate threads into page-locked blocks, next batch "
                        "inflating while the GPU counts) -> wd_count on host-mapped tiles -> counters on the host; wall clock, "
                        "max over ranks; flowcell_seconds = the same rate for %d tiles per GPU" % TILES_PER_LANE}
        if sweep:
            line["sweep_steps"] = sweep
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_tiles or 24 * cores
            planes = [p.array for p in pins]
            filts = [t.filt for t in tds]
            dt, wells, res = cpu_baseline_sample(planes, filts, centres, offs, idx, n_cpu, cores, args.hamming)
            ok = all(np.array_equal(res[k], counters[k]) for k in range(min(n_cpu, n_tiles)))
            line["cpu_baseline"] = {"value": n_cpu * N_TARGETS / dt, "unit": "targets/s", "cores": cores, "kind": "port",
                                    "wells_compared_per_s": wells / dt, "seconds": dt,
                                    "matches_gpu_counters": bool(ok),
                                    "sample": "%d tiles of the same lane (one per host thread), planes already gunzipped "
                                              "in RAM; C restatement of the reference (oracle/welldup_oracle.c)" % n_cpu,
                                    "python_reference_note": "the unmodified Python reference cannot travel to the GPU box; in the "
                                                             "authoring container it needs 9.35 s per tile on one core = 267 targets/s "
                                                             "(SURVEY section 6), 26x slower per core than this port"}
            line["counters_match_oracle"] = bool(ok)
        if not args.no_inflate and world == 1:
            line["host_inflate"] = host_inflate_sample(pins[0].array[0], os.cpu_count() or 1)
        report.write(json.dumps(line) + "\n")
        report.flush()
    if world > 1:
        dist.barrier()
        eng.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
