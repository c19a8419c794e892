"""Host-side mirrors of the reference interfaces (no GPU needed): target-file
parser, report writer, CLI surfaces, and that the C-ABI library loads and
exports every symbol include/welldup.h declares."""
import contextlib
import ctypes
import io
import json
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, ROOT
from oracle import ref_port as R
from well_duplicates_b200 import _lib, count_cli, prepare_cli, report
from well_duplicates_b200.targets import load_targets

REFT = os.path.join(GOLDEN, "ref_tests")
SMALL = os.path.join(REFT, "small.list")


# ---- target.py mirror: the reference's test/test_target.py, restated ----------------
def test_load_subset_and_limit():
    assert load_targets(SMALL, levels=2).levels == 2
    lim = load_targets(SMALL, limit=2)
    assert len(lim) == 2 and sum(1 for _ in lim) == 2


def test_get_all_indices():
    lim = load_targets(SMALL, levels=3, limit=2)
    assert set(lim.get_all_indices(0)) == {1998850, 3178500}
    assert set(lim.get_all_indices(1)) == set(map(int, (
        "1997278,1997279,1998849,1998851,2000420,2000421,3176929,3176930,3178499,3178501,3180071,3180072").split(",")))
    assert len(set(lim.get_all_indices(None))) == 38


def test_bad_files():
    with pytest.raises(ValueError):
        load_targets(os.path.join(REFT, "bad1.list"))
    with pytest.raises(AssertionError):
        load_targets(os.path.join(REFT, "bad2.list"))


def test_levels_targets_lookups():
    at = load_targets(SMALL)
    assert at.levels == 4 and at.get_target_by_centre(196654).get_levels() == 4
    assert len(at) == 7 and len(set(at.get_all_indices(0))) == 7
    res = at.get_from_index(196654)
    tgt = res[0][0]
    assert res == [(tgt, 0)] and tgt.get_centre() == 196654
    assert tgt.get_indices(1) == list(map(int, "195083,195084,196653,196655,198225,198226".split(",")))
    assert set(w for lev in range(4) for w in tgt.get_indices(lev)) == set(tgt.get_indices())
    assert len(set(at.get_all_indices())) == 213
    assert sorted(x[1] for x in at.get_from_index(1030466)) == [2, 2, 3]


def test_bad_add():
    at = load_targets(SMALL)
    with pytest.raises(Exception):
        at.add_target([(1, 2), (3, 4)])
    with pytest.raises(AssertionError):
        at.add_target([(111,), (112, 113, 114, 115)])
    sub = load_targets(SMALL, 2)
    sub.add_target([(111,), (112, 113, 114, 115)])
    with pytest.raises(Exception):
        sub.add_target([(111,), (112, 113, 114, 115)])


def test_to_csr_matches_oracle_parser():
    f = os.path.join(GOLDEN, "locs", "hex_small_n40_s13.list")
    for rings, limit in ((5, None), (3, 10), (1, 40)):
        at = load_targets(f, levels=rings + 1, limit=limit)
        centres, offs, idx = at.to_csr(rings)
        want = R.parse_target_file(f, levels=rings + 1, limit=limit)
        assert centres.tolist() == [t[0][0] for t in want]
        flat = [w for t in want for ring in t[1:] for w in ring]
        assert idx.tolist() == flat and offs[-1] == len(flat) and offs.size == len(want) * rings + 1
    with pytest.raises(IndexError):
        load_targets(f, levels=3).to_csr(5)


# ---- output_writer mirror: the reference's test/test_count_well_duplicates.py ------------
def test_output_writer_reference_cases():
    with open(os.path.join(REFT, "output_writer_cases.json")) as fh:
        cases = json.load(fh)
    for name, c in cases.items():
        dupl = {k: [[tuple(p) for p in t] for t in v] for k, v in c["lane_dupl"].items()}
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            report.output_writer(c["lane"], c["sample_size"], dupl, levels=c["levels"], verbose=c["verbose"])
        assert buf.getvalue() == c["printed"], name
        lines1 = buf.getvalue().rstrip("\n").split("\n")[:-4]
        lines2 = [re.sub(r"\s\s+", "\t", s) for s in c["expected"].lstrip().rstrip("\n").split("\n")]
        assert lines1 == lines2[c["sl"][0]:c["sl"][1]], name


def test_output_writer_zero_division_like_reference():
    with pytest.raises(ZeroDivisionError):
        report.output_writer(1, 4, {"1101": [[(0, 6), (0, 12)]]})
    with pytest.raises(ZeroDivisionError):
        report.output_writer(1, 4, {"1222": []}, levels=3)


def test_counters_round_trip_against_oracle():
    rng = np.random.default_rng(0)
    for _ in range(50):
        L = int(rng.integers(1, 6))
        tile = [[(int(rng.integers(0, 3)) * int(rng.random() < 0.3), int(rng.integers(1, 31))) for _ in range(L)]
                for _ in range(int(rng.integers(0, 30)))]
        row = report.counters_from_dupl(tile, L)
        n, wells, dups, hits, acco, acci = R.tile_counters(tile, L)
        assert row[0] == n and row[1::5].tolist() == wells and row[2::5].tolist() == dups
        assert row[3::5].tolist() == hits and row[4::5].tolist() == acco and row[5::5].tolist() == acci
        pt = np.array([[1] + [v for pair in t for v in pair] for t in tile] + [[0] * (1 + 2 * L)], dtype=np.int32)
        assert report.dupl_from_per_target(pt, L) == tile


# ---- CLI surfaces ---------------------------------------------------------------------------
def test_count_cli_flags_and_defaults():
    a = count_cli.parse_args(["-f", "x", "-s", "hiseq_x", "-r", "run"])
    assert (a.edit_distance, a.sample_size, a.level, a.start, a.end) == (2, 2500, 3, 50, 100)
    assert a.cycles is None and not a.hamming and not a.summary_only and not a.quiet and a.lane is None
    a = count_cli.parse_args("-f x -s 2228 -r run -e 1 -n 10 -l 5 -t 1101 -i 1,2 -x 3 -y 9 --cycles 0-5,8-14 --hamming -S -q".split())
    assert count_cli.parse_cycles(a) == [(0, 5), (8, 14)] and a.lane == "1,2" and a.tile_id == "1101"
    with pytest.raises(SystemExit):
        count_cli.parse_args(["-f", "x"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), pytest.raises(SystemExit):
        count_cli.parse_args(["--version"])
    assert buf.getvalue().strip() == "0.3"


def test_expected_tiles_equals_oracle():
    for st, t in (("hiseq_4000", None), ("hiseq_x", None), ("2488", None), ("2224", "1..[02468]"),
                  ("hiseq_x", "1101,1102"), ("junk", None), ("28", None)):
        assert count_cli.expected_tiles(st, t) == R.tile_list(st, t)
    with pytest.raises(AssertionError):
        count_cli.expected_tiles("hiseq_x", "9999")


def test_prepare_cli_surface(tmp_path):
    a = prepare_cli.parse_args(["-f", "s.locs"])
    assert a.seed is None and a.sample_size == 2500
    import random
    assert prepare_cli.get_random_array(1000, 5, 13) == R.sample_centres(1000, 5, 13)
    n, xy = prepare_cli.read_locs(os.path.join(GOLDEN, "locs", "hex_small.locs"))
    rn, rxy = R.read_locs(os.path.join(GOLDEN, "locs", "hex_small.locs"))
    assert n == rn and np.array_equal(xy, rxy)
    text = prepare_cli.format_targets([7, 9], np.array([0, 1, 3, 4, 5], np.uint32), np.array([1, 2, 3, 4, 5], np.uint32), levels=2)
    assert text == "7\n1\n2,3\n9\n4\n5\n"


# ---- the C-ABI library -------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "welldup.h")) as fh:
        declared = set(re.findall(r"^WD_API\s+[\w\s\*]+?\b(wd_\w+)\s*\(", fh.read(), flags=re.M))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().wd_abi_version() == _lib.ABI_VERSION == 2


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product refuses to run (it never falls back)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from well_duplicates_b200.engine import Engine
    with pytest.raises(_lib.CudaError, match="no CPU fallback"):
        Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "well_duplicates_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "ref_port" not in src and "c_port" not in src and "liboracle" not in src, f


# ---- binary target lists -----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ref_tests/small.list", "locs/hex_small_n40_s13.list", "locs/hex_tiny_n840_s3.list",
                                  "locs/negative_xy_n20_s2.list"])
def test_binary_target_list_round_trips_to_the_identical_text(name, tmp_path):
    import io

    from well_duplicates_b200 import targets as T
    src = os.path.join(GOLDEN, name)
    binary = str(tmp_path / "t.bin")
    T.text_to_binary(src, binary)
    assert T.is_binary_target_file(binary) and not T.is_binary_target_file(src)
    text = io.StringIO()
    T.binary_to_text(binary, text)
    with open(src) as fh:
        assert text.getvalue() == fh.read()
    # the mapped list answers like the parsed one: every levels / limit combination the CLI can ask for
    rings_on_file = T.load_targets(src).levels - 1
    for levels in [None] + list(range(1, rings_on_file + 2)):
        for limit in (None, 1, 5, 10 ** 6):
            a, b = T.load_targets(src, levels=levels, limit=limit), T.load_targets(binary, levels=levels, limit=limit)
            assert isinstance(b, T.BinaryTargets) and len(a) == len(b) and a.levels == b.levels
            for x, y in zip(a.to_csr(), b.to_csr()):
                assert np.array_equal(x, y)
            assert sorted(a.get_all_indices()) == sorted(b.get_all_indices())
            assert a.get_all_indices(0) == b.get_all_indices(0)
            if a.levels > 1:
                assert a.get_all_indices(1) == b.get_all_indices(1)
                assert [t.coords for t in a] == [t.coords for t in b]
                for r in range(a.levels):
                    for x, y in zip(a.to_csr(r), b.to_csr(r)):
                        assert np.array_equal(x, y)
            with pytest.raises(IndexError):
                b.to_csr(a.levels)


def test_binary_target_list_rejects_what_the_text_parser_rejects(tmp_path):
    from well_duplicates_b200 import targets as T
    path = str(tmp_path / "dup.bin")
    T.save_targets_binary(path, [7, 7], [0, 2, 4], [1, 2, 3, 4], 1)
    with pytest.raises(AssertionError):
        T.load_targets(path)                                     # duplicate centre (target.py:72)
    with pytest.raises(ValueError):
        T.save_targets_binary(path, [7, 8], [0, 2], [1, 2], 1)    # offsets do not match the target count
    with pytest.raises(ValueError):
        T.BinaryTargets(os.path.join(GOLDEN, "ref_tests", "small.list"))


def test_nccl_is_loaded_on_demand_and_hands_out_an_id():
    """wd_comm_unique_id needs no GPU: libnccl.so.2 is dlopen-ed on first use (csrc/wd_comm.cc); the 128 bytes are
    what rank 0 carries to the other ranks before wd_comm_init."""
    from well_duplicates_b200.engine import Engine
    try:
        a, b = Engine.comm_unique_id(), Engine.comm_unique_id()
    except _lib.CudaError as exc:
        pytest.skip("libnccl is not installed here: %s" % exc)
    assert len(a) == len(b) == _lib.COMM_ID_BYTES and a != b


def test_tuning_struct_matches_the_header_and_bench_sweep_items_map_onto_it():
    """wd_tuning (include/welldup.h) is 16 int32 whatever fields are named; every field the header declares is a
    field of the ctypes mirror and a keyword of Engine.set_tuning, and a bench.py --sweep-steps item parses to them."""
    import ctypes
    import inspect
    import re
    from bench_configs import parse_sweep
    from well_duplicates_b200 import _lib
    from well_duplicates_b200.engine import Engine
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "welldup.h")).read()
    body = header[header.index("typedef struct wd_tuning"):header.index("} wd_tuning;")]
    declared = [re.fullmatch(r"(\w+)(?:\[(\d+)\])?", item.strip()).groups()
                for decl in re.findall(r"int32_t\s+([^;]+);", body) for item in decl.split(",")]
    names = [n for n, dim in declared if not dim]
    assert sum(int(dim or 1) for _, dim in declared) == 16 and ctypes.sizeof(_lib.Tuning) == 64
    assert [f[0] for f in _lib.Tuning._fields_ if f[0] != "reserved"] == names
    assert set(names) <= set(inspect.signature(Engine.set_tuning).parameters)
    kw = parse_sweep("8,4,16 head_planes=3 targets_per_cta=64 ctas_per_sm=2")
    assert kw == {"step0": 8, "step1": 4, "centre_chunk": 16, "head_planes": 3, "targets_per_cta": 64, "ctas_per_sm": 2}
    assert set(kw) <= set(names)
