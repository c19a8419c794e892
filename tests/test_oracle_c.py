"""Pin the C restatement (oracle/welldup_oracle.c) to the golden outputs of the
reference and to the Python restatement."""
import json
import os

import numpy as np
import pytest

import fixture_inputs as fx
from helpers import GOLDEN, load_manifest, locs_path, parse_count_args
from oracle import c_port as CP
from oracle import ref_port as R

MAN = load_manifest()


@pytest.mark.parametrize("name", sorted(fx.LOCS_FIXTURES))
def test_pixels(name, tmp_path):
    _, xy = R.read_locs(locs_path(name, tmp_path))
    X, Y = R.locs_to_pixels(xy)
    x, y = CP.locs_to_pixels(xy)
    assert np.array_equal(x, X) and np.array_equal(y, Y)


@pytest.mark.parametrize("case", [c for c in MAN["prepare"] if c["n"] <= 60],
                         ids=lambda c: "%s_n%d_s%s" % (c["locs"], c["n"], c["seed"]))
def test_rings_match_reference_target_files(case, tmp_path):
    _, xy = R.read_locs(locs_path(case["locs"], tmp_path))
    X, Y = CP.locs_to_pixels(xy)
    if case["returncode"] != 0:
        with pytest.raises(RuntimeError):
            CP.ring_indexes(X, Y, R.sample_centres(xy.shape[0], case["n"], case["seed"])[0])
        return
    want = R.parse_target_file(os.path.join(GOLDEN, case["list"]))
    for tgt in want:
        assert CP.ring_indexes(X, Y, tgt[0][0]) == tgt[1:]


@pytest.mark.parametrize("case", [c for c in MAN["getseqs"]], ids=lambda c: c["name"])
def test_codes_match_reference(case):
    with open(os.path.join(GOLDEN, "getseqs", case["name"] + ".json")) as fh:
        want = json.load(fh)
    if "error" in want or case["start"] == case["end"]:
        pytest.skip("range / index errors are host logic")
    planes, kinds, filt, n = R.load_tile_planes(os.path.join(GOLDEN, case["run"]), case["lane"], case["tile"],
                                                case["start"], case["end"])
    keys = sorted(set(case["indices"]))
    codes, pf = CP.get_codes(planes, kinds, filt, keys)
    got = {str(k): [R.codes_to_str(codes[i]), bool(pf[i])] for i, k in enumerate(keys)}
    assert got == want["ok"]


def test_levenshtein_matches_python_port():
    rng = np.random.default_rng(11)
    for _ in range(500):
        n = int(rng.integers(1, 80))
        a = rng.integers(0, 5, n).astype(np.uint8)
        b = a.copy()
        for _ in range(int(rng.integers(0, 6))):
            b[int(rng.integers(0, n))] = rng.integers(0, 5)
        if rng.random() < 0.4:
            b = np.roll(b, int(rng.integers(-2, 3)))
        assert CP.levenshtein(a, b) == R.levenshtein(R.codes_to_str(a), R.codes_to_str(b))


@pytest.mark.parametrize("case", [c for c in MAN["count"] if c["returncode"] == 0 and c["name"] in (
    "lev_default", "hamming", "e0", "e4", "multirange", "limit10", "long75_e3", "cbcl_default", "cbcl_odd", "cbcl_late")],
    ids=lambda c: c["name"])
def test_count_tile_matches_python_port(case):
    o = parse_count_args(case["args"])
    run = os.path.join(GOLDEN, case["run"])
    targets = R.parse_target_file(os.path.join(GOLDEN, case["targets"]), levels=o["levels"] + 1, limit=o["limit"])
    centres = [t[0][0] for t in targets]
    offs, idx = [0], []
    for t in targets:
        for ring in t[1:]:
            idx.extend(ring)
            offs.append(len(idx))
    lane = o["lanes"].split(",")[0]
    for tile in R.tile_list(o["stype"], o["tiles"]):
        planes, kinds = [], []
        for s, e in o["ranges"]:
            p, k, filt, n = R.load_tile_planes(run, lane, tile, s, e)
            planes += p
            kinds += k
        pt, counters = CP.count_tile(planes, kinds, filt, centres, offs, idx, o["levels"], o["edit"], o["hamming"])
        seq_objs = [R.get_seqs_run(run, lane, tile, R.all_indices(targets), s, e) for s, e in o["ranges"]]
        want = R.count_tile(targets, seq_objs, o["levels"], o["edit"], o["hamming"])
        got = [[(int(r[1 + 2 * l]), int(r[2 + 2 * l])) for l in range(o["levels"])] for r in pt if r[0]]
        assert got == want
        n_t, wells, dups, hits, acco, acci = R.tile_counters(want, o["levels"])
        assert counters[0] == n_t
        assert counters[1::5].tolist() == wells and counters[2::5].tolist() == dups
        assert counters[3::5].tolist() == hits and counters[4::5].tolist() == acco and counters[5::5].tolist() == acci


def _exhaustive_manifest():
    with open(os.path.join(GOLDEN, "exhaustive", "manifest.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("case", _exhaustive_manifest(), ids=lambda c: c["name"])
def test_exhaustive_mode_matches_reference_report(case, tmp_path):
    """Every well a target: the oracle's counters, printed, equal what the
    unmodified reference printed (tests/golden/make_golden_exhaustive.py)."""
    from well_duplicates_b200 import report
    o = parse_count_args(case["args"])
    _, xy = R.read_locs(locs_path(case["locs"], tmp_path))
    X, Y = CP.locs_to_pixels(xy)
    rows = []
    for tile in case["tiles"]:
        planes, kinds = [], []
        for s, e in o["ranges"]:
            p, kd, filt, n = R.load_tile_planes(os.path.join(GOLDEN, "run_bcl"), case["lane"], tile, s, e)
            planes += p
            kinds += kd
        rows.append(CP.count_exhaustive(X, Y, planes, kinds, filt, o["levels"], o["edit"], o["hamming"]))
    with open(os.path.join(GOLDEN, "exhaustive", case["name"] + ".stdout")) as fh:
        want = fh.read()
    assert report.format_report(case["lane"], len(X), case["tiles"], rows, o["levels"], verbose=True) == want
