"""Whole-flowcell driver: every lane of a run in one command, tiles sharded over
the GPUs of one box.

Replaces what the reference does with one process per lane plus a text
concatenation (Snakefile.count_dups:143-160: rule count_well_dupl per lane, rule
summarize_all_lanes = ``tail`` of the per-lane files): the (lane, tile) pairs are
flattened in the reference's print order, rank r takes ordinals r, r+N, ...,
each rank counts its tiles on its own GPU, and the integer counter rows
``[Targets, (Wells, Dups, Hit, AccO, AccI) x levels]`` are combined by ONE
all-reduce (int64, sum) -- NCCL over NVLink on the GPUs, gloo in the CPU tests.
Every tile row is non-zero on exactly one rank, so the sum doubles as a gather
for the per-tile printout, and the per-lane totals come out of the same call.
Integer sums make the result independent of the rank count.

    python -m torch.distributed.run --nproc-per-node 8 -m well_duplicates_b200.flowcell \\
        -f targets.list -s 2224 -r RUN -l 5 --cycles 20-70 [-i 1,2,...] [-S]
"""
import os
import sys

import numpy as np

from . import count_cli
from .report import write_report


def plan(n_lanes, tiles_per_lane, rank, world):
    """Ordinals (lane-major, the reference's loop order) owned by ``rank``."""
    total = n_lanes * tiles_per_lane
    return np.arange(rank, total, world, dtype=np.int64)


def rows_layout(n_lanes, tiles_per_lane):
    """Row index of every tile and of every lane total in the exchange buffer."""
    total = n_lanes * tiles_per_lane
    return total, total + n_lanes


def exchange_host(local_rows, mine, n_lanes, tiles_per_lane, width, dist=None):
    """All-reduce of host rows (gloo / single process).  Returns the
    [tiles + lanes, width] int64 array every rank ends up with."""
    import torch
    n_tile_rows, n_rows = rows_layout(n_lanes, tiles_per_lane)
    buf = np.zeros((n_rows, width), dtype=np.int64)
    if len(mine):
        buf[mine] = local_rows
        np.add.at(buf, n_tile_rows + mine // tiles_per_lane, local_rows)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(buf)
        dist.all_reduce(t)
    return buf


def print_reports(stream, lanes, tiles, buf, sample_size, levels, verbose):
    """Per-lane reports, in lane order, from the exchanged rows."""
    tpl = len(tiles)
    order = np.argsort(np.array(tiles, dtype=object), kind="stable") if tiles else []
    for li, lane in enumerate(lanes):
        rows = buf[li * tpl:(li + 1) * tpl]
        write_report(stream, lane, sample_size, [tiles[k] for k in order], [rows[k] for k in order], levels, verbose)


def count_rank_tiles(eng, rd, stager, lanes, tiles, mine, wanted, levels, edit_distance, hamming, say=None,
                     publish=False, log_pairs=None):
    """Counter rows [len(mine), 1 + 5 * levels] of this rank's tiles (``mine`` = ordinals into lanes x tiles).
    Files -> page-locked planes on native threads one batch ahead of the GPU (staging.py); the planes stay in
    host memory and the fused kernel pulls the sectors it needs (wd_tile_map_host).  ``publish``: every batch's
    rows are also placed in the engine's exchange buffer on the device (K7, wd_publish_counters / _add), ready
    for ONE all-reduce after the last batch.  ``log_pairs`` = (say, centres, n_unique, n_ranges): this rank's
    tiles are also logged as count_well_duplicates.py does without -q (:211-222, :258-262), tile by tile."""
    from . import _lib
    from .staging import lane_batches
    rows = np.zeros((len(mine), 1 + 5 * levels), dtype=np.int64)
    names = ["%s/%s" % (lanes[int(o) // len(tiles)], tiles[int(o) % len(tiles)]) for o in mine]
    n_tile_rows, n_rows = rows_layout(len(lanes), len(tiles))

    def open_tile(name):
        lane, tile = name.split("/")
        return rd.get_tile(lane, tile)

    k = 0
    for got, staged in lane_batches(stager, open_tile, names, wanted):
        plane_of = stager.deliver(eng, staged, first_slot=0, zero_copy=True)
        order = [plane_of[c] for c in wanted]
        eng.count_async(0, len(got), order, edit_distance, hamming, mode=_lib.MODE_FUSED_LOG if log_pairs else _lib.MODE_FUSED)
        if publish:
            part = mine[k:k + len(got)]
            eng.publish_counters(part.astype(np.int32), (n_tile_rows + part // len(tiles)).astype(np.int32), n_rows, add=k > 0)
        rows[k:k + len(got)] = eng.count_fetch()[1]
        tile_logs = None
        if log_pairs:
            from .count_cli import tile_log_writers
            tile_logs = tile_log_writers(eng.dup_pairs(with_seqs=True), log_pairs[1], len(order))
        for j, name in enumerate(got):
            lane, tile = name.split("/")
            if say is not None:
                say("Reading tile %s in lane %s" % (tile, lane))
            if tile_logs is not None:
                log_pairs[0]("Got %i sequences from %i contiguous cycle ranges." % (log_pairs[2] * log_pairs[3], log_pairs[3]))
                tile_logs(j, log_pairs[0])
        k += len(got)
    if publish and k == 0:
        eng.publish_counters(np.zeros(0, np.int32), np.zeros(0, np.int32), n_rows)      # a rank without tiles
    return rows


def main(argv=None):
    import torch.distributed as dist

    from . import reader as bcl_direct_reader
    from .engine import Engine
    from .targets import load_targets

    args = count_cli.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    report_stream = sys.stdout
    if world > 1:
        # stdout is the report (a compatibility surface) but native libraries write there too
        # (NCCL prints its version banner on fd 1): keep a private handle on the real stdout
        # and point fd 1 at stderr for the rest of the process.
        sys.stdout.flush()
        report_stream = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        from .engine import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)          # staging memory next to this rank's GPU
        if not dist.is_initialized():
            # rendezvous only (the 128-byte NCCL id, the final barrier): the counters travel by the library's
            # own ncclAllReduce on device memory (wd_comm_init / wd_allreduce_i64)
            dist.init_process_group("gloo")
    # every rank logs its own tiles (stderr of a multi-rank run interleaves; each tile's lines stay together)
    say = (lambda *a: None) if args.quiet else count_cli.log

    lanes = args.lane.split(",") if args.lane else [str(x) for x in range(1, 9)]
    tiles = count_cli.expected_tiles(args.stype, args.tile_id)
    cycles = count_cli.parse_cycles(args)
    wanted = [c for s, e in cycles for c in range(s, e)]
    targets = load_targets(filename=args.coord_file, levels=args.level + 1, limit=args.sample_size)
    eng = Engine(local)
    centres, level_offsets, idx = targets.to_csr(args.level)
    eng.load_targets(centres, level_offsets, idx, args.level)
    if world > 1:
        uid = [Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
    rd = bcl_direct_reader.BCLReader(args.run, engine=eng)

    mine = plan(len(lanes), len(tiles), rank, world)
    width = 1 + 5 * args.level
    from .staging import Stager
    stager = Stager(threads=max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))),
                    cbcl_cache=rd._cbcl_cache)
    log_pairs = None if args.quiet else (say, centres, len(targets.get_all_indices()), len(cycles))
    count_rank_tiles(eng, rd, stager, lanes, tiles, mine, wanted, args.level, args.edit_distance, args.hamming, say,
                     publish=True, log_pairs=log_pairs)
    stager.close()
    # K7 has put every batch's rows into the exchange buffer on the device: one ncclAllReduce(int64, sum) over NVLink
    eng.allreduce_published()
    buf = eng.published_fetch(rows_layout(len(lanes), len(tiles))[1] * width).reshape(-1, width)
    if rank == 0:
        print_reports(report_stream, lanes, tiles, buf, len(targets), args.level, verbose=not args.summary_only)
        report_stream.flush()
    if world > 1:
        dist.barrier()
        eng.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
