"""Command line of count_well_duplicates.py (reference :272-317) on the GPU.

Same flags, defaults, stdout and stderr.  The host walks lanes and tiles, reads
and inflates the files, and prints; the gather-decode, the distance tests and
the counter reduction run in the CUDA library."""
import re
import sys
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np

from . import reader as bcl_direct_reader
from .report import write_report
from .targets import load_targets

__VERSION__ = 0.3
HISEQ_4000 = "hiseq_4000"
HISEQ_X = "hiseq_x"


def log(msg):
    # one write per line: the ranks of a multi-GPU run share stderr, and print() would hand text and newline over separately
    sys.stderr.write(str(msg) + "\n")


def parse_args(argv=None):
    p = ArgumentParser(description="Counts well duplicates on a patterned flowcell without mapping: reads "
                                   "within LEVEL rings of each sampled well are compared with it.",
                       formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("-f", "--coord_file", dest="coord_file", help="target list from prepare_cluster_indexes")
    p.add_argument("-e", "--edit_distance", dest="edit_distance", type=int, default=2,
                   help="largest distance that still counts as a duplicate")
    p.add_argument("-n", "--sample_size", dest="sample_size", type=int, default=2500,
                   help="number of targets to take from the list")
    p.add_argument("-l", "--level", dest="level", type=int, default=3, help="rings around each centre to test, max = 5")
    p.add_argument("-s", "--stype", dest="stype", required=True,
                   help="%s, %s, or the highest tile number (e.g. 2228) from which swaths and tiles are inferred" % (HISEQ_4000, HISEQ_X))
    p.add_argument("-r", "--run", dest="run", required=True, help="run folder (the one holding Data/)")
    p.add_argument("-t", "--tile", dest="tile_id", type=str, help="comma-separated tiles or regexes, e.g. 1..[02468]")
    p.add_argument("-i", "--lane", dest="lane", type=str, help="comma-separated lanes, 1-8")
    p.add_argument("-x", "--start", dest="start", type=int, default=50, help="first cycle of the compared substring")
    p.add_argument("-y", "--end", dest="end", type=int, default=100, help="cycle after the last one compared")
    p.add_argument("--cycles", help="list of cycle ranges, e.g. 10-50,100-120; overrides -x/-y")
    p.add_argument("--hamming", action="store_true", help="Hamming distance instead of Levenshtein")
    p.add_argument("-S", "--summary-only", action="store_true", help="print only the per-lane summary")
    p.add_argument("-q", "--quiet", action="store_true", help="no log output")
    p.add_argument("--version", action="version", version=str(__VERSION__))
    # extension (not in the reference): what "prepare_cluster_indexes.py -n <every well>" followed by this
    # command would report, without the multi-gigabyte target file and its ~86 h of preparation per tile
    p.add_argument("--exhaustive-locs", metavar="S_LOCS", help="every well of each tile is a target; rings come "
                   "straight from this .locs file (replaces -f; -n is ignored)")
    args = p.parse_args(argv)
    if not args.coord_file and not args.exhaustive_locs:
        p.error("the following arguments are required: -f/--coord_file")
    return args


def expected_tiles(stype, tile_id=None):
    """Tile names of one lane (count_well_duplicates.py:164-191)."""
    max_tile, max_swath = 24, 22
    if stype == HISEQ_4000:
        max_tile = 28
    else:
        try:
            max_tile = int(stype) % 100
            max_swath = int(stype) // 100 or 22
        except ValueError:
            pass
    tiles = ["%d%d%02d" % (surface, swath, t)
             for surface in range(1, max_swath // 10 + 1)
             for swath in range(1, max_swath % 10 + 1)
             for t in range(1, max_tile + 1)]
    if tile_id:
        keep = set()
        for pat in tile_id.split(","):
            hits = [t for t in tiles if re.match("^" + pat + "$", t)]
            # the reference means to assert here but trips over an undefined name (:189)
            assert hits, "%s matches no tile identifiers for a %s" % (pat, stype)
            keep.update(hits)
        tiles = sorted(keep)
    return tiles


def parse_cycles(args):
    if args.cycles:
        return [(int(s), int(e)) for r in args.cycles.split(",") for s, e in (r.split("-"),)]
    return [(args.start, args.end)]


def tile_log_writers(logged, centres, seq_len):
    """``logged`` = Engine.dup_pairs(with_seqs=True) of a batch (or None) -> f(tile of the batch, say) that writes
    that tile's duplicate pairs as the reference does (count_well_duplicates.py:258-262)."""
    if logged is None or not len(logged[0]):
        return lambda k, say: None
    pairs, codes = logged
    seqs = bcl_direct_reader.codes_to_strings(codes.reshape(-1, seq_len))
    with_rows, at = np.unique(pairs[:, 0], return_index=True)       # rows are sorted by tile
    first = dict(zip(with_rows.tolist(), at.tolist()))

    def write(k, say):
        i = first.get(k, len(pairs))
        while i < len(pairs) and pairs[i, 0] == k:
            _, t_ord, well, dist = pairs[i].tolist()
            say("center seq at {:>07}: {}".format(int(centres[t_ord]), seqs[2 * i]))
            say("well seq at   {:>07}: {}".format(well, seqs[2 * i + 1]))
            say("edit distance: {}".format(dist))
            i += 1
    return write


def main_exhaustive(args):
    """--exhaustive-locs: wd_count_exhaustive per tile, same report."""
    from .prepare_cli import read_locs
    from .report import write_report
    from .staging import lane_batches
    say = (lambda *a: None) if args.quiet else log
    lanes = args.lane.split(",") if args.lane else range(1, 8 + 1)
    tiles = expected_tiles(args.stype, args.tile_id)
    wanted = [c for s, e in parse_cycles(args) for c in range(s, e)]
    bcl_reader = bcl_direct_reader.BCLReader(args.run)
    eng = bcl_reader.engine
    _, xy = read_locs(args.exhaustive_locs)
    eng.load_locs(xy)
    stager = bcl_direct_reader.default_stager(bcl_reader._cbcl_cache)
    for lane in lanes:
        rows = []
        # the dense pack reads every byte of every plane: planes are copied to HBM (DMA from the
        # pinned block) while the next tile inflates
        for names, batch in lane_batches(stager, lambda t: bcl_reader.get_tile(lane, t), tiles, wanted, per_batch=1,
                                         announce=lambda t: say("Reading tile %s in lane %s" % (t, lane))):
            plane_of = stager.deliver(eng, batch, first_slot=0, zero_copy=False)
            rows.append(eng.count_exhaustive(0, [plane_of[c] for c in wanted], args.level, args.edit_distance, args.hamming))
        order = sorted(range(len(tiles)), key=lambda k: tiles[k])
        write_report(sys.stdout, lane, xy.shape[0], [tiles[k] for k in order], [rows[k] for k in order], args.level,
                     verbose=not args.summary_only)


def main(argv=None):
    args = parse_args(argv)
    if args.exhaustive_locs:
        return main_exhaustive(args)
    say = (lambda *a: None) if args.quiet else log
    lanes = args.lane.split(",") if args.lane else range(1, 8 + 1)
    tiles = expected_tiles(args.stype, args.tile_id)
    cycles = parse_cycles(args)
    targets = load_targets(filename=args.coord_file, levels=args.level + 1, limit=args.sample_size)
    bcl_reader = bcl_direct_reader.BCLReader(args.run)
    if len(targets) == 0 and tiles:
        # an empty list: the reference opens the first tile and fails in get_seqs (bcl_direct_reader.py:186)
        lane0 = str(lanes[0] if args.lane else 1)
        say("Reading tile %s in lane %s" % (tiles[0], lane0))
        bcl_reader.get_tile(lane0, tiles[0])
        raise IndexError("list index out of range")
    eng = bcl_reader.engine
    centres, level_offsets, idx = targets.to_csr(args.level)
    eng.load_targets(centres, level_offsets, idx, args.level)
    n_unique = len(targets.get_all_indices())
    wanted = [c for s, e in cycles for c in range(s, e)]
    from . import _lib
    from .staging import lane_batches
    stager = bcl_direct_reader.default_stager(bcl_reader._cbcl_cache)

    # Files -> page-locked planes on native threads, one batch ahead of the GPU (staging.py), across
    # lane boundaries: the first tiles of the next lane inflate while this lane is finished and printed.
    # Many tiles per launch; the planes stay in host memory and the fused kernel pulls the sectors it
    # needs.  Without -q the same kernel also logs every duplicate pair (count_well_duplicates.py:258-262)
    # and the log is printed tile by tile, in the reference's order, once the batch has been counted.
    lanes = [str(lane) for lane in lanes]
    walk = ["%s/%s" % (lane, t) for lane in lanes for t in tiles]
    lane_rows = {lane: {} for lane in lanes}      # counter rows of the device reduction
    todo = {lane: len(tiles) for lane in lanes}

    def open_tile(name):
        lane, t = name.split("/")
        return bcl_reader.get_tile(lane, t)

    def announce(name):
        lane, t = name.split("/")
        say("Reading tile %s in lane %s" % (t, lane))

    def finish(lane):
        names = sorted(lane_rows[lane])
        # output_writer infers the level count from the first tile with a valid target: none -> no level lines
        levels = args.level if any(int(lane_rows[lane][t][0]) for t in names) else 0
        write_report(sys.stdout, lane, len(targets), names, [lane_rows[lane][t][:1 + 5 * levels] for t in names], levels,
                     verbose=not args.summary_only)

    mode = _lib.MODE_FUSED if args.quiet else _lib.MODE_FUSED_LOG
    for names, staged in lane_batches(stager, open_tile, walk, wanted, on_error=announce):
        plane_of = stager.deliver(eng, staged, first_slot=0, zero_copy=True)
        order = [plane_of[c] for c in wanted]
        try:
            _, counters = eng.count(0, len(names), order, args.edit_distance, args.hamming, mode=mode, per_target=False)
        except IndexError:
            announce(names[0])         # a well beyond the tile: the reference fails on the first tile of this size
            raise
        tile_logs = tile_log_writers(None if args.quiet else eng.dup_pairs(with_seqs=True), centres, len(order))
        for k, name in enumerate(names):
            lane, tname = name.split("/")
            announce(name)
            say("Got %i sequences from %i contiguous cycle ranges." % (n_unique * len(cycles), len(cycles)))
            tile_logs(k, say)
            lane_rows[lane][tname] = counters[k]
            todo[lane] -= 1
            if todo[lane] == 0:
                finish(lane)
    if not tiles:
        for lane in lanes:
            finish(lane)


if __name__ == "__main__":
    main()
