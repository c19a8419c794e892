"""Pin the CPU oracle (oracle/ref_port.py) to outputs of the unmodified
reference (tests/golden/, made by tests/golden/make_golden.py) and to the
reference's own test expectations (test/test_target.py,
test/test_count_well_duplicates.py)."""
import json
import os
import re

import numpy as np
import pytest

import fixture_inputs as fx
from helpers import GOLDEN, load_manifest, locs_path, parse_count_args
from oracle import ref_port as R

MAN = load_manifest()


# ---------------------------------------------------------------- stage 1 --
@pytest.mark.parametrize("case", MAN["prepare"], ids=lambda c: "%s_n%d_s%s" % (c["locs"], c["n"], c["seed"]))
def test_prepare_matches_reference(case, tmp_path):
    path = locs_path(case["locs"], tmp_path)
    with open(os.path.join(GOLDEN, case["list"])) as fh:
        want = fh.read()
    if case["returncode"] != 0:
        assert case["error"] == "RuntimeError" and want == ""
        with pytest.raises(RuntimeError):
            R.prepare_cluster_indexes(path, case["n"], case["seed"])
        return
    assert R.prepare_cluster_indexes(path, case["n"], case["seed"]) == want


def test_prepare_loop_equals_vectorised():
    path = locs_path("hex_small")
    n, xy = R.read_locs(path)
    X, Y = R.locs_to_pixels(xy)
    for c in (0, 47, 48, 1500, 3071, 1234):
        assert R.ring_indexes_loop(X, Y, c) == R.ring_indexes(X, Y, c)


def test_integer_ring_rule_equals_float_rule():
    """The d^2 thresholds used on the GPU equal the reference's float64 sqrt rule."""
    thr = [d * d for d in R.MAX_DISTS]
    import math
    for dx in range(0, 131):
        for dy in range(0, 131):
            d2 = dx * dx + dy * dy
            dist = math.sqrt(d2)
            for lev in range(5):
                assert (R.MAX_DISTS[lev] < dist <= R.MAX_DISTS[lev + 1]) == (thr[lev] < d2 <= thr[lev + 1])


# ------------------------------------------------------------ target file --
REFT = os.path.join(GOLDEN, "ref_tests")


def test_target_file_reference_expectations():
    """test/test_target.py:37-82, :94-107 restated on the oracle parser."""
    small = os.path.join(REFT, "small.list")
    t = R.parse_target_file(small)
    assert len(t) == 7 and all(len(x) == 4 for x in t)
    assert len(R.parse_target_file(small, levels=2)[0]) == 2
    lim = R.parse_target_file(small, levels=3, limit=2)
    assert len(lim) == 2
    assert {x[0][0] for x in lim} == {1998850, 3178500}
    assert {i for x in lim for i in x[1]} == set(map(int, (
        "1997278,1997279,1998849,1998851,2000420,2000421,"
        "3176929,3176930,3178499,3178501,3180071,3180072").split(",")))
    assert len(set(R.all_indices(lim))) == 38
    assert len(set(R.all_indices(t))) == 213          # test_target.py:113-116
    by_centre = {x[0][0]: x for x in t}
    assert by_centre[196654][1] == list(map(int, "195083,195084,196653,196655,198225,198226".split(",")))


def test_target_file_bad_inputs():
    """test/test_target.py:84-90."""
    with pytest.raises(ValueError):
        R.parse_target_file(os.path.join(REFT, "bad1.list"))
    with pytest.raises(AssertionError):
        R.parse_target_file(os.path.join(REFT, "bad2.list"))


# ---------------------------------------------------------------- stage 2 --
@pytest.mark.parametrize("case", MAN["getseqs"], ids=lambda c: c["name"])
def test_get_seqs_matches_reference(case):
    with open(os.path.join(GOLDEN, "getseqs", case["name"] + ".json")) as fh:
        want = json.load(fh)
    run = os.path.join(GOLDEN, case["run"])
    if "error" in want:
        assert want["error"] == "IndexError"
        with pytest.raises(IndexError):
            R.get_seqs_run(run, case["lane"], case["tile"], case["indices"], case["start"], case["end"])
        return
    got = R.get_seqs_run(run, case["lane"], case["tile"], case["indices"], case["start"], case["end"])
    assert {str(k): [v[0], v[1]] for k, v in got.items()} == want["ok"]


def test_filter_offsets_example():
    """Docstring example of cbcl_read.py:186-190: 000110101 -> [-1,-1,-1,0,1,-1,2,-1,3]."""
    assert R.filter_offsets([0, 0, 0, 1, 1, 0, 1, 0, 1]).tolist() == [-1, -1, -1, 0, 1, -1, 2, -1, 3]


# ---------------------------------------------------------------- stage 3 --
def test_distance_known_answers():
    assert R.levenshtein("kitten", "sitting") == 3
    assert R.levenshtein("ACGT", "ACGT") == 0
    assert R.levenshtein("ACGTACGT", "CGTACGTA") == 2
    assert R.hamming("ACGTACGT", "CGTACGTA") == 8
    assert R.levenshtein("NNNN", "NNNN") == 0 and R.hamming("ANNA", "NNNN") == 2


def test_distance_properties():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 40))
        a = "".join("ACGTN"[i] for i in rng.integers(0, 5, n))
        b = "".join("ACGTN"[i] for i in rng.integers(0, 5, n))
        c = "".join("ACGTN"[i] for i in rng.integers(0, 5, n))
        assert R.levenshtein(a, b) == R.levenshtein(b, a)
        assert R.levenshtein(a, b) <= R.hamming(a, b)
        assert R.levenshtein(a, c) <= R.levenshtein(a, b) + R.levenshtein(b, c)


@pytest.mark.parametrize("case", MAN["count"], ids=lambda c: c["name"])
def test_count_matches_reference(case):
    o = parse_count_args(case["args"])
    with open(os.path.join(GOLDEN, "count", case["name"] + ".stdout")) as fh:
        want = fh.read()
    run = os.path.join(GOLDEN, case["run"])
    tf = os.path.join(GOLDEN, case["targets"])
    tiles = R.tile_list(o["stype"], o["tiles"])
    lanes = o["lanes"].split(",")
    got = ""
    log = []
    try:
        for lane in lanes:
            got += R.count_run(run, tf, lane, tiles, o["levels"], o["ranges"], sample_size=o["limit"],
                               edit_distance=o["edit"], use_hamming=o["hamming"],
                               verbose=not o["summary"], log=log)
    except (ZeroDivisionError, FileNotFoundError, RuntimeError) as exc:
        # the reference died the same way (manifest "raises"), after printing the lanes in front of the failure
        assert case["returncode"] != 0 and type(exc).__name__ == (case.get("raises") or "ZeroDivisionError")
        if not isinstance(exc, ZeroDivisionError):
            assert got == want
        return
    assert case["returncode"] == 0
    assert got == want
    # the duplicate-pair log on stderr (count_well_duplicates.py:258-262)
    with open(os.path.join(GOLDEN, "count", case["name"] + ".stderr")) as fh:
        err = fh.read()
    if not o["quiet"]:
        pairs = re.findall(r"center seq at (\d+): (\S*)\nwell seq at +(\d+): (\S*)\nedit distance: (\d+)\n", err)
        assert [(int(a), b, int(c), d, int(e)) for a, b, c, d, e in pairs] == log


def test_output_writer_reference_cases():
    """test/test_count_well_duplicates.py:96-148: the expected strings there
    predate the 4 trailing lines output_writer now prints; both are checked."""
    with open(os.path.join(REFT, "output_writer_cases.json")) as fh:
        cases = json.load(fh)
    for name, c in cases.items():
        dupl = {k: [[tuple(p) for p in t] for t in v] for k, v in c["lane_dupl"].items()}
        got = R.report_text(c["lane"], c["sample_size"], dupl, levels=c["levels"], verbose=bool(c["verbose"]))
        assert got == c["printed"], name
        lines1 = got.rstrip("\n").split("\n")[:-4]
        lines2 = [re.sub(r"\s\s+", "\t", s) for s in c["expected"].lstrip().rstrip("\n").split("\n")]
        lines2 = lines2[c["sl"][0]:c["sl"][1]]
        assert lines1 == lines2, name


def test_tile_list_shapes():
    assert len(R.tile_list("hiseq_4000")) == 112
    assert len(R.tile_list("hiseq_x")) == 96
    assert len(R.tile_list("2224")) == 96
    assert len(R.tile_list("2488")) == 704
    assert R.tile_list("hiseq_x", "1..[02468]")[:3] == ["1102", "1104", "1106"]
