"""Deterministic inputs shared by tests/golden/make_golden.py (which feeds them
to the unmodified reference) and the tests (which feed them to the oracle and
to the CUDA path).  Coordinates are built with integer arithmetic only, so the
files are identical on every machine."""
import numpy as np

from well_duplicates_b200 import synth

SMALL_ROW_LEN = 48
SMALL_ROWS = 64
SMALL_WELLS = SMALL_ROW_LEN * SMALL_ROWS     # 3072
SMALL_CYCLES = 14


def hex_small():
    X, Y = synth.hex_lattice(SMALL_WELLS, SMALL_ROW_LEN)
    return synth.xy_to_locs_floats(X, Y)


def hex_tiny():
    X, Y = synth.hex_lattice(30 * 28, 28)
    return synth.xy_to_locs_floats(X, Y)


def hex_shuffled():
    """Same lattice as hex_small but the wells are stored in a scrambled order,
    so index order is unrelated to position (the reference makes no raster
    assumption; neither may a spatial index)."""
    X, Y = synth.hex_lattice(SMALL_WELLS, SMALL_ROW_LEN)
    perm = (np.arange(SMALL_WELLS, dtype=np.int64) * 1237 + 11) % SMALL_WELLS
    return synth.xy_to_locs_floats(X[perm], Y[perm])


def window_cm():
    """Column-major 4000 x 11 grid (x pitch 20 px, y pitch 3 px): wells 5
    columns apart differ by 20000 +- a few in index, so the reference's
    asymmetric scan window [c-20000, c+20001] decides membership."""
    rows, cols = 4000, 11
    idx = np.arange(rows * cols, dtype=np.int64)
    col, row = idx // rows, idx % rows
    X = (1100 + 20 * col).astype(np.int32)
    Y = (1050 + 3 * row).astype(np.int32)
    return synth.xy_to_locs_floats(X, Y)


def sparse3():
    """Three wells far apart: every ring is empty -> the reference aborts."""
    X = np.array([1500, 5000, 9000], dtype=np.int32)
    Y = np.array([1500, 5000, 9000], dtype=np.int32)
    return synth.xy_to_locs_floats(X, Y)


def negative_xy():
    """Lattice translated so pixel coordinates go negative (x < -100.05 in the
    file): int() truncation toward zero differs from floor there."""
    X, Y = synth.hex_lattice(40 * 30, 30, x0=-400, y0=-300)
    xy = np.empty((X.size, 2), dtype=np.float32)
    xy[:, 0] = (X.astype(np.float64) - 1000.0) / 10.0
    xy[:, 1] = (Y.astype(np.float64) - 1000.0) / 10.0
    return xy


LOCS_FIXTURES = {
    "hex_small": hex_small,
    "hex_tiny": hex_tiny,
    "hex_shuffled": hex_shuffled,
    "window_cm": window_cm,
    "sparse3": sparse3,
    "negative_xy": negative_xy,
}
LOCS_NOT_COMMITTED = ["window_cm"]

# (locs fixture, sample size, seed)
PREPARE_CASES = [
    ("hex_small", 40, 13),
    ("hex_small", 25, 7),
    ("hex_tiny", 840, 3),          # every well a target (exhaustive mode, small)
    ("hex_shuffled", 30, 11),
    ("window_cm", 12, 5),
    ("sparse3", 2, 1),             # RuntimeError, empty stdout
    ("negative_xy", 20, 2),
]

_BCL = ["-s", "hiseq_x", "-i", "1", "-t", "1101,1102"]
_CBCL = ["-s", "2488", "-i", "1"]
# (name, run dir, extra argv)
COUNT_CASES = [
    ("lev_default", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14"]),
    ("hamming", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "--hamming"]),
    ("e0", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-e", "0"]),
    ("e1", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-e", "1"]),
    ("e4", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-e", "4"]),
    ("e5_hamming", "run_bcl", _BCL + ["-l", "4", "--cycles", "0-14", "-e", "5", "--hamming"]),
    ("multirange", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-5,8-14"]),
    ("xy_l3", "run_bcl", _BCL + ["-x", "2", "-y", "12"]),
    ("summary", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-S"]),
    ("limit10", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-n", "10"]),
    ("two_lanes", "run_bcl", ["-s", "hiseq_x", "-i", "1,2", "-t", "1101", "-l", "5", "--cycles", "0-14"]),
    ("regex_tiles", "run_bcl", ["-s", "hiseq_x", "-i", "1", "-t", "110[13],1102", "-l", "2", "--cycles", "0-14"]),
    ("quiet", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14", "-q"]),
    ("long70", "run_bcl", _BCL + ["-l", "5", "--cycles", "0-14,0-14,0-14,0-14,0-14"]),
    ("long75_e3", "run_bcl", _BCL + ["-l", "5", "-e", "3", "--cycles", "0-14,2-13,1-14,0-12,3-14,0-14"]),
    ("long140_hamming", "run_bcl", _BCL + ["-l", "3", "--hamming", "--cycles", ",".join(["0-14"] * 10)]),
    ("zero_hits", "run_bcl", ["-s", "hiseq_x", "-i", "1", "-t", "1103", "-l", "5", "--cycles", "0-14",
                              "-e", "0", "--hamming"]),
    ("cbcl_default", "run_cbcl", _CBCL + ["-t", "1101,2101", "-l", "5", "--cycles", "0-14"]),
    ("cbcl_hamming", "run_cbcl", _CBCL + ["-t", "1101,2101", "-l", "5", "--cycles", "0-14", "--hamming"]),
    ("cbcl_odd", "run_cbcl", _CBCL + ["-t", "1102", "-l", "5", "--cycles", "3-12"]),
    ("cbcl_early", "run_cbcl", _CBCL + ["-t", "1101", "-l", "5", "--cycles", "0-6"]),
    ("cbcl_late", "run_cbcl", _CBCL + ["-t", "1101", "-l", "5", "--cycles", "6-14"]),
    # lane 3 does not exist: the reference prints the reports of lanes 1 and 2, logs "Reading tile" for the
    # tile it cannot open and dies with FileNotFoundError
    ("missing_lane", "run_bcl", ["-s", "hiseq_x", "-i", "1,2,3", "-t", "1101", "-l", "5", "--cycles", "0-14"]),
    # ... and a tile without a .filter file in the middle of a lane: RuntimeError after the tiles in front of it
    ("missing_tile", "run_bcl", ["-s", "hiseq_x", "-i", "1", "-t", "110[1-4]", "-l", "5", "--cycles", "0-14"]),
]

# the per-lane report files a workflow run leaves behind (Snakefile.count_and_push:174-181), for the
# all-lanes summary (`tail`, :166-172) and the two wiki formatters that parse it (:183-197)
WIKI_LANES = [("1", ["-s", "hiseq_x", "-i", "1", "-t", "1101,1102", "-l", "5", "--cycles", "0-14"]),
              ("2", ["-s", "hiseq_x", "-i", "2", "-t", "1101", "-l", "5", "--cycles", "0-14", "-S"])]


def getseqs_cases():
    some = [0, 1, 2, 5, 100, 1500, 3071, 47, 48, 2999, 1234, 777]
    many = list(range(0, 3072, 61))
    return [
        # name, run, lane, tile, indices, start, end
        ("bcl_some", "run_bcl", 1, 1101, some, 0, None),
        ("bcl_window", "run_bcl", 1, 1101, some, 3, 9),
        ("bcl_few", "run_bcl", 1, 1102, [7, 9, 3071], 0, None),
        ("bcl_many", "run_bcl", 1, 1102, many, 1, 13),
        ("bcl_dupidx", "run_bcl", 1, 1101, [5, 5, 6, 5, 2000, 6] * 3, 0, 4),
        ("bcl_lane_str", "run_bcl", "L002", 1101, some, 0, None),
        ("bcl_empty_range", "run_bcl", 1, 1101, some, 5, 5),
        ("bcl_oob", "run_bcl", 1, 1101, [0, 3072] + many, 0, None),
        ("bcl_neg", "run_bcl", 1, 1101, [-1, 4] + many, 0, None),
        ("cbcl_some", "run_cbcl", 1, 1101, some, 0, None),
        ("cbcl_early", "run_cbcl", 1, 1101, many, 0, 6),
        ("cbcl_late", "run_cbcl", 1, 1101, many, 6, 14),
        ("cbcl_odd", "run_cbcl", 1, 1102, some + [3072, 3070], 0, None),
        ("cbcl_surface2", "run_cbcl", 1, 2101, many, 2, 11),
    ]
