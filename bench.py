#!/usr/bin/env python3
"""Benchmark of the well_duplicates hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (per GPU; weak scaling, N GPUs = N lanes): one synthetic HiSeq-4000
lane of 96 tiles x 4 309 650 wells, 2500 sampled targets out to ring 5, a
50-cycle substring from BCL byte planes, default (Levenshtein, e = 2) compare.
A step is one pass of the hot path (fused gather-decode + compare + counter
reduction, then the lane-counter all-reduce when N > 1) over all of the rank's
tiles.  `value` has the planes resident in HBM.  `e2e` starts every step from
planes in pinned HOST memory and ends with the counters on the host, through
the C ABI: the product's staging (wd_tile_map_host) leaves the planes where the
inflate step wrote them and the kernel pulls the 32-byte sectors it needs across
PCIe; `e2e_staged` is the same step with every plane copied to HBM first
(wd_tile_put_bcl), for comparison.  Gunzip is outside all of them
(BASELINE.json north_star).

`--impl reference` times the CPU restatement of the reference (oracle/, C port,
all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WELLS = 4309650
ROW_LEN = 1571
N_CYCLES = 50
N_TARGETS = 2500
LEVELS = 5
EDIT = 2
TILES_PER_LANE = 96
SEED = 20261018
METRIC = "targets/sec (with wells compared/sec) on synthetic HiSeq 4000 lanes, 96 tiles x 2500 targets per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=TILES_PER_LANE, help="tiles per GPU")
    ap.add_argument("--distinct-tiles", type=int, default=12,
                    help="distinct synthetic tiles kept in pinned host memory (reused round-robin for the tile slots)")
    ap.add_argument("--mode", type=int, default=0, help="0 fused kernel, 1 two-pass kernels")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-mode", default="both", choices=["staged", "zerocopy", "both"],
                    help="staged: every plane copied to HBM through wd_tile_put_bcl; zerocopy: planes stay in pinned "
                         "host memory (wd_tile_map_host) and the kernel pulls the sectors it needs over PCIe")
    ap.add_argument("--cpu-tiles", type=int, default=0, help="tiles in the cpu_baseline sample (0 = 24 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inflate", action="store_true", help="skip the host gunzip measurement (reported beside the metric)")
    ap.add_argument("--hamming", action="store_true")
    ap.add_argument("--zc-blocks", type=int, default=0,
                    help="distinct pinned host blocks for the zero-copy e2e (0 = one per tile slot if the host can pin them)")
    ap.add_argument("--sweep-steps", default="", help="early-exit schedules to time in this process, e.g. '8,4;8,2;6,2' "
                                                      "(sets WELLDUP_STEPS; resident planes, and zero-copy when --e2e-mode asks)")
    ap.add_argument("--l2-fetch", type=int, default=0, help="override cudaLimitMaxL2FetchGranularity (32/64/128)")
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

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6 or not (t0 - 0.05 <= ts <= t1 + 0.15):
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
                "samples": len(sm)}


def algorithmic_bytes(centres, offs, idx, filt, fused=True):
    """Bytes one tile's launch has to move ('Algorithmic bytes')."""
    valid = (filt[centres] & 1).astype(bool)
    lens = np.diff(offs.astype(np.int64)).reshape(-1, LEVELS).sum(axis=1)
    t_off = np.concatenate([[0], np.cumsum(lens)])
    if fused:
        keep = np.repeat(valid, lens)
        wells = np.concatenate([idx[keep], centres[valid]])
    else:
        wells = np.concatenate([idx, centres])
    sectors = np.unique(wells >> 5).size
    slots = int(wells.size)
    plane_bytes = sectors * 32 * N_CYCLES
    filt_bytes = np.unique(centres >> 5).size * 32
    index_bytes = slots * 5 + centres.size * 8            # slot_well u32 + slot_level u8, tgt_off
    out_bytes = (1 + 5 * LEVELS) * 8
    if fused:
        return plane_bytes + filt_bytes + index_bytes + out_bytes, sectors, slots
    packed = slots * 32
    return plane_bytes + np.unique(wells >> 5).size * 32 + index_bytes + 2 * packed + out_bytes, sectors, slots


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def load_traffic(key):
    """Per-launch byte counts taken from committed ncu captures (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get(key)
    return None


def cpu_baseline_sample(planes_by_tile, filts, centres, offs, idx, n_tiles, threads, hamming):
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
        res = list(pool.map(one, range(n_tiles)))
    dt = time.perf_counter() - t0
    wells = int(sum(int(r[1::5].sum()) for r in res))
    return dt, wells, res


def host_inflate_sample(plane, threads, seconds=4.0):
    """Gunzip stays on the host (BASELINE.json north_star) and is reported beside the metric: the product's
    inflate (wd_inflate_batch, csrc/wd_inflate.cc -- what staging.Stager runs) on .bcl.gz-shaped members
    (one compressed plane of the benchmark tile) on all host threads, and zlib -- what the reference's
    gzip.open().read() runs -- on the same members."""
    import zlib
    from well_duplicates_b200 import _lib
    raw = struct_header(plane.size) + plane.tobytes()
    comp = zlib.compress(raw, 1)
    comp = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + comp[2:-4] + struct_pack_tail(raw)
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


def struct_header(n):
    import struct
    return struct.pack("<I", n)


def struct_pack_tail(raw):
    import struct
    import zlib
    return struct.pack("<II", zlib.crc32(raw) & 0xffffffff, len(raw) & 0xffffffff)


def run_reference(args, rank, world):
    """The reference's CPU path (C port of it) on the box's host cores."""
    if rank != 0:
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
            "wells_per_tile": N_WELLS, "kernel": "fused" if args.mode == 0 else "two-pass",
            "distinct_tiles": args.distinct_tiles, "seed": SEED,
            "l2": "inputs (%.1f GB of planes per GPU) are far larger than L2; no flush needed" % (
                per_gpu_tiles * N_WELLS * N_CYCLES / 1e9),
            "parallelism": "tiles sharded ordinal % n_gpus; one int64 all-reduce of the counter rows per step"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries ONE JSON line: native libraries write there too (NCCL prints its version banner on
    # fd 1), so keep a private handle on the real stdout and point fd 1 at stderr for the rest of the run
    sys.stdout.flush()
    report = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import Engine, PinnedArray

    torch.cuda.set_device(local)
    if world > 1:
        from well_duplicates_b200.engine import bind_to_gpu_numa_node
        bound = None if os.environ.get("WELLDUP_NO_NUMA_BIND") else bind_to_gpu_numa_node(local)
        sys.stderr.write("rank %d: GPU %d, host cores %s\n" % (rank, local, "unchanged" if bound is None else
                                                               "%s (next to %s)" % (_ranges(bound[1]), bound[0])))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    l2_prev = eng.set_l2_fetch_granularity(args.l2_fetch) if args.l2_fetch else None
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)

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

    class _Cai:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def step(fetch):
        eng.count_async(0, n_tiles, order, EDIT, args.hamming, mode=args.mode, per_target=False)
        if world > 1:
            ptr, n = eng.publish_counters(tile_row, lane_row, n_rows)
            t = torch.as_tensor(_Cai(ptr, n), device=torch.device("cuda", local))
            dist.all_reduce(t)                       # ncclAllReduce(int64, sum) over NVLink
            if fetch:
                return t.cpu().numpy().reshape(n_rows, -1)
            return None
        if fetch:
            return eng.count_fetch()[1]
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step(False)
        counters = step(True)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.3)
        l0 = eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_w0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            step(False)
        ev1.record(stream)
        barrier()
        t_w1 = time.time()
        launches = eng.launch_count() - l0
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(t_w0, t_w1)

        def timed(fn, n):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(n):
                fn()
            b.record(stream)
            barrier()
            return a.elapsed_time(b) / n

        sweep = {}
        schedules = [x for x in args.sweep_steps.split(";") if x]
        def set_schedule(sch):
            # "first,later[,centre_chunk]" and/or "NAME=value" tunables, space separated; None restores the defaults
            for k in [k for k in os.environ if k.startswith("WELLDUP_")]:
                os.environ.pop(k)
            for item in (sch or "").split():
                if "=" in item:
                    k, v = item.split("=", 1)
                    os.environ["WELLDUP_" + k] = v
                else:
                    f = item.split(",")
                    os.environ["WELLDUP_STEPS"] = ",".join(f[:2])
                    if len(f) > 2:
                        os.environ["WELLDUP_CENTRE_CHUNK"] = f[2]

        for sch in schedules:
            set_schedule(sch)
            step(False)
            sweep[sch] = {"resident_ms": timed(lambda: step(False), args.steps)}
        set_schedule(None)

        # ---- e2e, zero-copy flavour: planes stay in pinned host memory ------------------
        zc_ms = None
        if args.e2e_steps > 0 and args.e2e_mode in ("zerocopy", "both"):
            # one distinct pinned block per tile slot, so that no slot can be served from lines another
            # slot brought into L2 -- unless the host cannot pin that much (then the D blocks are shared)
            n_zc = n_tiles if args.zc_blocks <= 0 else max(D, min(n_tiles, args.zc_blocks))
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            need = (n_zc - D) * (N_CYCLES + 1) * N_WELLS * local_world
            try:
                import psutil
                if args.zc_blocks <= 0 and psutil.virtual_memory().available < 1.25 * need + (8 << 30):
                    n_zc = D
            except ImportError:
                pass
            zc = list(pins) + [PinnedArray((N_CYCLES, N_WELLS)) for _ in range(n_zc - D)]
            zc_filt = list(filt_pins) + [PinnedArray((N_WELLS,)) for _ in range(n_zc - D)]
            with ThreadPoolExecutor(max_workers=8) as pool:
                list(pool.map(lambda s: np.copyto(zc[s].array, pins[s % D].array), range(D, n_zc)))
            for s in range(D, n_zc):
                zc_filt[s].array[:] = filt_pins[s % D].array

            def map_tiles():
                for s in range(n_tiles):
                    eng.tile_map_host(s, N_WELLS, zc[s % n_zc].array, pinned_filter=zc_filt[s % n_zc].array)

            map_tiles()
            zc_counters = step(True)
            zc_ok = bool(np.array_equal(zc_counters, counters))
            barrier()
            ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev4.record(stream)
            for _ in range(args.e2e_steps):
                map_tiles()
                res = step(True)
            ev5.record(stream)
            barrier()
            zc_ms = ev4.elapsed_time(ev5) / args.e2e_steps
            d2h_per_step = int(res.size * 8)
            zc_dma_bytes = eng.last_count_h2d_bytes()

            def zc_step():
                map_tiles()
                step(True)
            for sch in schedules:
                set_schedule(sch)
                zc_step()
                sweep[sch]["zero_copy_ms"] = timed(zc_step, args.e2e_steps)
            set_schedule(None)
            for z in zc[D:]:
                z.free()

        # ---- e2e: host planes -> C ABI -> counters on the host, every step -------------
        e2e_ms = None
        if args.e2e_steps > 0 and args.e2e_mode in ("staged", "both"):
            push_tiles()
            step(True)
            barrier()
            ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev2.record(stream)
            for _ in range(args.e2e_steps):
                push_tiles()
                res = step(True)
            ev3.record(stream)
            barrier()
            e2e_ms = ev2.elapsed_time(ev3) / args.e2e_steps
            d2h_per_step = int(res.size * 8)

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms or 0.0, zc_ms or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        e2e_ms = float(t[1]) if e2e_ms is not None else None
        zc_ms = float(t[2]) if zc_ms is not None else None

    ms_per_step = ms / args.steps
    targets_per_step = total_tiles * N_TARGETS
    if world > 1:
        wells_per_step = int(counters[:total_tiles, 1::5].sum())
    else:
        wells_per_step = int(counters[:, 1::5].sum())
    value = targets_per_step / (ms_per_step / 1e3)

    if rank == 0:
        peak, peak_kind = load_peaks()
        per_tile = [algorithmic_bytes(centres, offs, idx, tds[k].filt, fused=(args.mode == 0)) for k in range(D)]
        alg_bytes = sum(per_tile[s % D][0] for s in range(n_tiles))
        achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
        traffic = load_traffic("fused" if args.mode == 0 else "two_pass")
        line = {
            "metric": METRIC, "value": value, "unit": "targets/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64 bit-planes", "data": "synthetic",
            "wells_compared_per_s": wells_per_step / (ms_per_step / 1e3),
            "config": workload_config(args, n_tiles),
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json (%s)" % peak_kind,
                         "kernel": "fused_count_kernel" if args.mode == 0 else "gather_pack_kernel+compare_count_kernel",
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "algorithmic_bytes_note": "SURVEY 8(d) / DESIGN 4.3: every 32-byte sector that holds a well of a "
                                                   "pass-filter target, in all %d planes, + filter, index and counter bytes. "
                                                   "The kernel stops reading a well as soon as its prefix proves dist > e, so "
                                                   "it moves fewer bytes than that (traffic); dram_achieved / dram_frac are "
                                                   "the rate at which it really drives HBM" % N_CYCLES,
                         "dram_achieved": None if not traffic else traffic * (n_tiles / TILES_PER_LANE) / (ms_per_step / 1e3) / 1e9,
                         "dram_frac": None if not traffic else traffic * (n_tiles / TILES_PER_LANE) / (ms_per_step / 1e3) / 1e9 / peak,
                         "distinct_32B_sectors_per_plane_per_tile": int(np.mean([p[1] for p in per_tile])),
                         "note": "duration = CUDA events around %d launches on the launching stream; at N>1 it also "
                                 "covers the publish kernel and the all-reduce" % args.steps},
        }
        staged = None if e2e_ms is None else {
            "value": targets_per_step / (e2e_ms / 1e3), "unit": "targets/s", "ms_per_step": e2e_ms,
            "staging": "wd_tile_put_bcl: every plane copied from pinned host memory to HBM, then counted",
            "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": d2h_per_step,
            "h2d_gb_per_s": h2d_per_step / (e2e_ms / 1e3) / 1e9}
        if zc_ms is not None:
            pulled = load_traffic("zero_copy_pcie_read_bytes")
            line["e2e"] = {
                "value": targets_per_step / (zc_ms / 1e3), "unit": "targets/s", "ms_per_step": zc_ms,
                "staging": "wd_tile_map_host: planes and filters stay in pinned host memory (%.1f GB per step and GPU); "
                           "wd_count copies the planes of the first 2 compared cycles to HBM by DMA, tile group after "
                           "tile group, while the counting kernel reads the sectors it needs of the later planes "
                           "across PCIe" % (n_tiles * (N_CYCLES + 1) * N_WELLS / 1e9),
                "h2d_bytes_per_step": int(zc_dma_bytes + (0 if pulled is None else pulled * n_tiles / TILES_PER_LANE)),
                "h2d_dma_bytes_per_step": int(zc_dma_bytes),
                "h2d_pulled_bytes_per_step": None if pulled is None else int(pulled * n_tiles / TILES_PER_LANE),
                "h2d_bytes_note": "DMA bytes are counted by the library; pulled bytes = pcie__read_bytes of the step's "
                                  "kernels in the committed ncu capture (profiles/traffic.json)",
                "host_bytes_mapped_per_step": int(n_tiles * (N_CYCLES + 1) * N_WELLS),
                "distinct_host_blocks": n_zc,
                "d2h_bytes_per_step": d2h_per_step, "counters_equal_staged_run": zc_ok}
            if staged is not None:
                line["e2e_staged"] = staged
        else:
            line["e2e"] = staged
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
                                              "in RAM; C restatement of the reference (oracle/welldup_oracle.c)" % n_cpu}
        if not args.no_inflate and world == 1:
            line["host_inflate"] = host_inflate_sample(pins[0].array[0], os.cpu_count() or 1)
        report.write(json.dumps(line) + "\n")
        report.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
