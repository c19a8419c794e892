"""Per-tile / LaneSummary printout.

``output_writer`` keeps the signature and the exact text of the reference's
output_writer (count_well_duplicates.py:27-153).  The integer work is done on
counter rows ``[Targets, (Wells, Dups, Hit, AccO, AccI) x levels]`` -- the same
rows the CUDA reduction (K6) produces -- so ``format_report`` can print
straight from device results."""
import sys

import numpy as np

TALLY = 0
LENGTH = 1


def counters_from_dupl(tile_counts, levels):
    """One tile's ``[[(tally, length)] * levels] * valid_targets`` -> counter row."""
    row = np.zeros(1 + 5 * levels, dtype=np.int64)
    if len(tile_counts) == 0 or levels == 0:
        row[0] = len(tile_counts)
        return row
    a = np.array([[pair for pair in targ[:levels]] for targ in tile_counts], dtype=np.int64)  # [t, L, 2]
    tally, length = a[:, :, TALLY], a[:, :, LENGTH]
    hit = tally > 0
    acco = np.maximum.accumulate(hit, axis=1)                       # hit at this ring or further in
    acci = np.maximum.accumulate(hit[:, ::-1], axis=1)[:, ::-1]     # ... or further out
    row[0] = a.shape[0]
    row[1::5] = length.sum(axis=0)
    row[2::5] = tally.sum(axis=0)
    row[3::5] = hit.sum(axis=0)
    row[4::5] = acco.sum(axis=0)
    row[5::5] = acci.sum(axis=0)
    return row


def dupl_from_per_target(per_target, levels):
    """Device per-target rows ``[valid, (dups, wells) x levels]`` -> the
    reference's lane_dupl entry for that tile."""
    out = []
    for r in per_target[per_target[:, 0] != 0]:
        out.append([(int(r[1 + 2 * l]), int(r[2 + 2 * l])) for l in range(levels)])
    return out


def write_report(stream, lane, sample_size, tiles, counters, levels, verbose=False):
    """Writes the report for tiles (names, already in print order) with counter
    rows.  Lines go out as they are produced, so a lane without any hit prints
    its per-tile lines and then raises ZeroDivisionError exactly where the
    reference does (count_well_duplicates.py:115-117)."""
    def emit(line):
        stream.write(line + "\n")

    tot = [0] * (1 + 5 * levels)
    for tile, row in zip(tiles, counters):
        row = [int(v) for v in row]
        if verbose:
            emit("Lane: %s\tTile: %s\tTargets: %i/%i" % (lane, tile, row[0], sample_size))
            for lev in range(levels):
                w, d, h, o, i = row[1 + 5 * lev: 6 + 5 * lev]
                emit("Level: %i\tWells: %i\tDups: %i\tHit: %i\tAccO: %i\tAccI: %i" % (lev + 1, w, d, h, o, i))
        tot = [a + b for a, b in zip(tot, row[: len(tot)])]
    targets = tot[0]
    if levels:
        hits_any = tot[5]                               # AccI at level 1 = targets with a hit anywhere
        dups_all = sum(tot[2::5])
        peds = hits_any * (1 - hits_any / (dups_all + hits_any)) / targets
        peds2 = hits_any * (1 - hits_any / (2 * dups_all)) / targets
    else:
        hits_any = peds = peds2 = 0
    emit("LaneSummary: %s\tTiles: %i\tTargets: %i/%i" % (lane, len(tiles), targets, sample_size * len(tiles)))
    for lev in range(levels):
        w, d, h, o, i = tot[1 + 5 * lev: 6 + 5 * lev]
        emit("Level: %i\tWells: %i\tDups: %i (%.5f)\tHit: %i (%.5f)\tAccO: %i (%.5f)\tAccI: %i (%.5f)" % (
            lev + 1, w, d, d / w, h, h / targets, o, o / targets, i, i / targets))
    raw = hits_any / targets if hits_any else 0.0
    emit("")
    emit("Overall duplication (Acc/Targets): {:.2%}".format(raw))
    emit("Picard-equivalent duplication v1:  {:.2%}".format(peds))
    emit("Picard-equivalent duplication v2:  {:.2%}".format(peds2))


def format_report(lane, sample_size, tiles, counters, levels, verbose=False):
    import io
    buf = io.StringIO()
    write_report(buf, lane, sample_size, tiles, counters, levels, verbose)
    return buf.getvalue()


def output_writer(lane, sample_size, lane_dupl, levels=0, verbose=False):
    """Drop-in for the reference function of the same name: ``lane_dupl[tile]``
    is a list (one entry per target whose centre passed the filter) of
    ``[(TALLY, LENGTH)] * levels``."""
    if not levels:
        for tile_counts in lane_dupl.values():
            if len(tile_counts) > 0:
                levels = len(tile_counts[0])
                break
    tiles = sorted(lane_dupl.keys())
    rows = [counters_from_dupl(lane_dupl[t], levels) for t in tiles]
    write_report(sys.stdout, lane, sample_size, tiles, rows, levels, verbose)
