"""Parity of the CUDA path (through the C ABI and the Python mirrors of the
reference interfaces) with the golden outputs of the unmodified reference and
with the CPU oracle on seeded inputs.  Integer / byte work: everything is
compared bit-exactly."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

import fixture_inputs as fx
from helpers import GOLDEN, load_manifest, locs_path, parse_count_args

pytestmark = pytest.mark.gpu
MAN = load_manifest()


@pytest.fixture(scope="module")
def eng():
    from well_duplicates_b200.reader import default_engine
    return default_engine()


@pytest.fixture(scope="module")
def oracle():
    from oracle import c_port, ref_port
    return ref_port, c_port


# ---------------------------------------------------------------- stage 1 --
@pytest.mark.parametrize("name", sorted(fx.LOCS_FIXTURES))
def test_k0_pixels(eng, oracle, name, tmp_path):
    R, CP = oracle
    _, xy = R.read_locs(locs_path(name, tmp_path))
    eng.load_locs(xy)
    x, y = eng.pixels()
    X, Y = R.locs_to_pixels(xy)
    assert np.array_equal(x, X) and np.array_equal(y, Y)


@pytest.mark.parametrize("case", MAN["prepare"], ids=lambda c: "%s_n%d_s%s" % (c["locs"], c["n"], c["seed"]))
def test_prepare_cli_byte_identical(case, tmp_path, capsys):
    """prepare_cluster_indexes drop-in: stdout equals the reference's target file, stderr its chatter (seed,
    sample, the byte offset logged for every scan; up to the traceback when the reference dies on an empty ring);
    --binary writes the same list as arrays that convert back to the byte-identical text."""
    import io

    from well_duplicates_b200 import prepare_cli, targets
    path = locs_path(case["locs"], tmp_path)
    with open(os.path.join(GOLDEN, case["list"])) as fh:
        want = fh.read()
    with open(os.path.join(GOLDEN, case["list"][:-len(".list")] + ".stderr")) as fh:
        want_err = fh.read()
    argv = ["-f", path, "-n", str(case["n"])] + (["-s", str(case["seed"])] if case["seed"] is not None else [])
    if case["returncode"] != 0:
        with pytest.raises(RuntimeError, match="Got no wells"):
            prepare_cli.main(argv)
        got = capsys.readouterr()
        assert got.out == "" and got.err == want_err
        return
    prepare_cli.main(argv)
    got = capsys.readouterr()
    assert got.out == want and got.err == want_err
    binary = str(tmp_path / "targets.bin")
    prepare_cli.main(argv + ["--binary", binary])
    assert capsys.readouterr().out == ""
    text = io.StringIO()
    targets.binary_to_text(binary, text)
    assert text.getvalue() == want


def test_ring_query_every_well_vs_oracle(eng, oracle, tmp_path):
    """Every well of the shuffled lattice and a slice of the window lattice
    against the oracle's scan (window rule, ordering, level bins)."""
    R, CP = oracle
    for name, centres in (("hex_shuffled", None), ("window_cm", list(range(19000, 19040)) + list(range(43900, 44000)) + [0, 1, 3999, 4000])):
        _, xy = R.read_locs(locs_path(name, tmp_path))
        X, Y = CP.locs_to_pixels(xy)
        centres = list(range(xy.shape[0])) if centres is None else centres
        eng.load_locs(xy)
        offs, idx = eng.ring_query(centres, 5)
        woffs, widx = CP.rings_csr(X, Y, centres)
        assert np.array_equal(offs, woffs) and np.array_equal(idx, widx)


def test_ring_query_fewer_levels_and_bad_centre(eng, oracle):
    R, CP = oracle
    _, xy = R.read_locs(locs_path("hex_small"))
    X, Y = CP.locs_to_pixels(xy)
    eng.load_locs(xy)
    offs, idx = eng.ring_query([5, 1500], 2)
    woffs, widx = CP.rings_csr(X, Y, [5, 1500], 2)
    assert np.array_equal(offs, woffs) and np.array_equal(idx, widx)
    with pytest.raises(IndexError):
        eng.ring_query([xy.shape[0]], 5)


# ---------------------------------------------------------------- stage 2 --
@pytest.mark.parametrize("case", MAN["getseqs"], ids=lambda c: c["name"])
def test_get_seqs_reader_api(case):
    """BCLReader(...).get_tile(...).get_seqs(...) equals the reference's dict."""
    from well_duplicates_b200.reader import BCLReader
    with open(os.path.join(GOLDEN, "getseqs", case["name"] + ".json")) as fh:
        want = json.load(fh)
    tile = BCLReader(os.path.join(GOLDEN, case["run"])).get_tile(case["lane"], case["tile"])
    if "error" in want:
        with pytest.raises(IndexError):
            tile.get_seqs(case["indices"], case["start"], case["end"])
        return
    got = tile.get_seqs(case["indices"], case["start"], case["end"])
    assert {str(k): [v[0], v[1]] for k, v in got.items()} == want["ok"]
    assert all(isinstance(k, int) and isinstance(v[1], bool) for k, v in got.items())


def test_abi_rejects_out_of_range_wells(eng):
    """The C entry point itself raises IndexError, not only the Python mirror."""
    eng.tile_begin(0, 100, 1)
    eng.tile_put_filter(0, np.ones(100, np.uint8))
    eng.tile_put_bcl(0, 0, np.full(100, 5, np.uint8))
    with pytest.raises(IndexError, match="out of range"):
        eng.get_seqs(0, [3, 100], [0])
    with pytest.raises(IndexError, match="negative"):
        eng.get_seqs(0, [-2, 5], [0])
    with pytest.raises(AssertionError):
        eng.tile_put_bcl(0, 0, np.zeros(99, np.uint8))      # header mismatch, bcl_direct_reader.py:338


@pytest.mark.parametrize("run,lane,tile", [("run_bcl", 1, 1101), ("run_cbcl", 1, 1102)])
def test_k3_filter_offsets(oracle, run, lane, tile):
    from well_duplicates_b200.reader import BCLReader
    R, CP = oracle
    t = BCLReader(os.path.join(GOLDEN, run)).get_tile(lane, tile)
    got = t._get_filter_offsets()
    want = R.filter_offsets(t.read_filter())
    assert got == want.tolist() and t.passing_wells == int((want >= 0).sum())


def test_k3_rank_large_random(eng, oracle):
    R, CP = oracle
    rng = np.random.default_rng(1)
    for n in (1, 63, 64, 65, 4097, 1000003):
        filt = rng.integers(0, 4, n, dtype=np.uint8)
        eng.tile_begin(1, n, 0)
        eng.tile_put_filter(1, filt)
        off, passing = eng.filter_offsets(1)
        woff, wpass = CP.filter_offsets(filt)
        assert np.array_equal(off, woff) and passing == wpass


# ---------------------------------------------------------------- stage 3 --
@pytest.mark.parametrize("case", MAN["count"], ids=lambda c: c["name"])
def test_count_cli_matches_reference(case):
    """count_well_duplicates drop-in: stdout (report) and stderr (log with every
    duplicate pair) equal the reference's."""
    from helpers import golden_count_output, run_count_case
    from well_duplicates_b200 import count_cli
    want_out, want_err = golden_count_output(case)
    out, err = run_count_case(count_cli.main, case)
    assert out == want_out
    assert err == want_err


@pytest.mark.parametrize("name", ["lev_default", "limit10", "xy_l3", "cbcl_default"])
def test_count_cli_reads_binary_target_lists(name, tmp_path):
    """-f with the binary form of the same list (targets.py): same report, same log."""
    from helpers import golden_count_output, run_count_case
    from well_duplicates_b200 import count_cli, targets
    case = dict([c for c in MAN["count"] if c["name"] == name][0])
    binary = str(tmp_path / "targets.bin")
    targets.text_to_binary(os.path.join(GOLDEN, case["targets"]), binary)
    case["targets"] = binary
    assert run_count_case(count_cli.main, case) == golden_count_output(case)


def test_report_text_feeds_the_wiki_formatters_gpu(tmp_path):
    """Lane reports of the drop-in CLI -> summarize_all_lanes (`tail`) -> the reference's wiki formatters
    (restated in tests/wiki_formatters.py, pinned to the unmodified scripts' output in tests/golden/wiki)."""
    import wiki_formatters as W
    from well_duplicates_b200 import count_cli
    from well_duplicates_b200 import workflow as wf
    wiki = os.path.join(GOLDEN, "wiki")
    lane_files = []
    for lane, args in MAN["wiki"]["lanes"]:
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            count_cli.main(["-f", os.path.join(GOLDEN, MAN["wiki"]["targets"]), "-r", os.path.join(GOLDEN, MAN["wiki"]["run"]), "-q"] + args)
        path = tmp_path / ("40targets_lane%s.txt" % lane)
        path.write_text(out.getvalue())
        lane_files.append(str(path))
    for tag, extra in (("all_lanes", 0), ("all_lanes_plus4", 3)):
        text = wf.summarize_all_lanes(lane_files, levels=5, extra=extra, names=[os.path.basename(p) for p in lane_files])
        with open(os.path.join(wiki, "40targets_%s.txt" % tag)) as fh:
            assert text == fh.read()
        for fn, ext in ((W.to_wiki, "wiki"), (W.to_wiki2, "wiki2.html")):
            with open(os.path.join(wiki, "40targets_%s.%s" % (tag, ext))) as fh:
                assert fn(text) == fh.read()


def _load_case(eng, R, case, o):
    """Stage the golden run of a count case; returns (planes order, tiles, targets)."""
    from well_duplicates_b200.reader import BCLReader
    from well_duplicates_b200.targets import load_targets
    targets = load_targets(os.path.join(GOLDEN, case["targets"]), levels=o["levels"] + 1, limit=o["limit"])
    centres, offs, idx = targets.to_csr(o["levels"])
    eng.load_targets(centres, offs, idx, o["levels"])
    rd = BCLReader(os.path.join(GOLDEN, case["run"]), engine=eng)
    wanted = [c for s, e in o["ranges"] for c in range(s, e)]
    tiles = R.tile_list(o["stype"], o["tiles"])
    lane = o["lanes"].split(",")[0]
    for k, t in enumerate(tiles):
        plane_of = rd.get_tile(lane, t).stage(k, wanted)
    return [plane_of[c] for c in wanted], tiles, (centres, offs, idx), lane


@pytest.mark.parametrize("name", ["lev_default", "hamming", "e4", "multirange", "long75_e3", "long140_hamming",
                                  "cbcl_default", "cbcl_odd"])
def test_fused_two_pass_and_oracle_agree(eng, oracle, name):
    """Both kernel flavours, batched over the case's tiles, against the C oracle:
    per-target (dups, wells) rows and the tile counters."""
    R, CP = oracle
    case = [c for c in MAN["count"] if c["name"] == name][0]
    o = parse_count_args(case["args"])
    order, tiles, (centres, offs, idx), lane = _load_case(eng, R, case, o)
    pt0, c0 = eng.count(0, len(tiles), order, o["edit"], o["hamming"], mode=0)
    pt1, c1 = eng.count(0, len(tiles), order, o["edit"], o["hamming"], mode=1)
    assert np.array_equal(pt0, pt1) and np.array_equal(c0, c1)
    run = os.path.join(GOLDEN, case["run"])
    for k, t in enumerate(tiles):
        planes, kinds = [], []
        for s, e in o["ranges"]:
            p, kd, filt, n = R.load_tile_planes(run, lane, t, s, e)
            planes += p
            kinds += kd
        wpt, wc = CP.count_tile(planes, kinds, filt, centres, offs, idx, o["levels"], o["edit"], o["hamming"])
        assert np.array_equal(pt0[k], wpt), (name, t)
        assert np.array_equal(c0[k], wc), (name, t)


def _synthetic_tile(seed, n_wells, row_len, n_cycles, n_targets, **kw):
    from oracle import c_port as CP
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(seed)
    X, Y = synth.hex_lattice(n_wells, row_len)
    td = synth.make_tile(rng, n_wells, n_cycles, row_len, **kw)
    centres = rng.choice(n_wells, size=n_targets, replace=False).astype(np.uint32)
    return X, Y, td, centres


@pytest.mark.parametrize("e,ham", [(2, False), (2, True), (0, False), (3, False), (6, False), (60, False)])
def test_medium_tile_vs_oracle(eng, oracle, e, ham):
    """300 x 400 lattice, 50 cycles, 1500 targets: stage 1 on the GPU feeds
    stages 2+3; everything checked against the C oracle."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    X, Y, td, centres = _synthetic_tile(7, 120000, 400, 50, 1500, dup_rate=0.3, shift_share=0.4, nocall_rate=0.003)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    woffs, widx = CP.rings_csr(X, Y, centres)
    assert np.array_equal(offs, woffs) and np.array_equal(idx, widx)
    eng.load_targets(centres, offs, idx, 5)
    eng.tile_begin(0, td.n_wells, td.n_cycles)
    eng.tile_put_filter(0, td.filt)
    for c in range(td.n_cycles):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(td.n_cycles))
    wpt, wc = CP.count_tile([td.planes[c] for c in order], ["bcl"] * len(order), td.filt, centres, offs, idx, 5, e, ham)
    for mode in (0, 1):
        pt, cnt = eng.count(0, 1, order, e, ham, mode=mode)
        assert np.array_equal(pt[0], wpt) and np.array_equal(cnt[0], wc)
    if e >= 2:
        assert wc[2::5].sum() > 50      # the case does contain duplicates


def test_medium_cbcl_mixed_vs_oracle(eng, oracle):
    """NovaSeq-style planes: first cycles store every well, later ones only PF
    wells; odd PF count; three tiles in one launch."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(99)
    n, row_len, ncyc, split = 50001, 250, 40, 17
    X, Y = synth.hex_lattice(n, row_len)
    centres = rng.choice(n, size=600, replace=False).astype(np.uint32)
    woffs, widx = CP.rings_csr(X, Y, centres)
    eng.load_targets(centres, woffs, widx, 5)
    tiles = []
    for k in range(3):
        td = synth.make_tile(rng, n, ncyc, row_len, pf_rate=0.6, dup_rate=0.2, shift_share=0.3, nocall_rate=0.02)
        pfmask = (td.filt & 1).astype(bool)
        eng.tile_begin(k, n, ncyc)
        eng.tile_put_filter(k, td.filt)
        planes, kinds = [], []
        for c in range(ncyc):
            nib = synth.bcl_to_nibbles(td.planes[c])
            excl = c >= split
            if excl:
                nib = nib[pfmask]
            packed = synth.pack_nibbles(nib)
            eng.tile_put_cbcl(k, c, packed, nib.size, excl)
            planes.append(packed)
            kinds.append("cbcl_excl" if excl else "cbcl")
        tiles.append((planes, kinds, td.filt))
    order = list(range(3, ncyc))
    pt0, c0 = eng.count(0, 3, order, 2, False, mode=0)
    pt1, c1 = eng.count(0, 3, order, 2, False, mode=1)
    assert np.array_equal(pt0, pt1) and np.array_equal(c0, c1)
    for k, (planes, kinds, filt) in enumerate(tiles):
        wpt, wc = CP.count_tile([planes[c] for c in order], [kinds[c] for c in order], filt, centres, woffs, widx, 5, 2, False)
        assert np.array_equal(pt0[k], wpt) and np.array_equal(c0[k], wc)
    # a block whose cluster count disagrees with the filter is refused (cbcl_read.py:130-131)
    eng.tile_put_cbcl(0, ncyc - 1, tiles[0][0][ncyc - 1], int((tiles[0][2] & 1).sum()) - 1, True)
    with pytest.raises(AssertionError):
        eng.count(0, 1, order, 2, False, mode=0)


def test_zero_copy_planes_equal_staged(eng, oracle):
    """wd_tile_map_host: planes left in pinned host memory and read in place by
    the kernels give what staged planes give (BCL, and CBCL with both block
    kinds), for both kernel flavours and for get_seqs."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import PinnedArray
    rng = np.random.default_rng(5)
    n, row_len, ncyc, split = 30001, 200, 30, 11
    X, Y = synth.hex_lattice(n, row_len)
    centres = rng.choice(n, size=400, replace=False).astype(np.uint32)
    woffs, widx = CP.rings_csr(X, Y, centres)
    eng.load_targets(centres, woffs, widx, 5)
    td = synth.make_tile(rng, n, ncyc, row_len, pf_rate=0.7, dup_rate=0.25, shift_share=0.3, nocall_rate=0.01)
    order = list(range(ncyc))
    # BCL
    pin = PinnedArray((ncyc, n + 7))
    pin.array[:, :n] = td.planes
    eng.tile_map_host(0, n, pin.array)
    eng.tile_put_filter(0, td.filt)
    wpt, wc = CP.count_tile([td.planes[c] for c in order], ["bcl"] * ncyc, td.filt, centres, woffs, widx, 5, 2, False)
    for mode in (0, 1):
        pt, cnt = eng.count(0, 1, order, 2, False, mode=mode)
        assert np.array_equal(pt[0], wpt) and np.array_equal(cnt[0], wc)
    codes, pf = eng.get_seqs(0, widx[:500], order)
    wcodes, wpf = CP.get_codes([td.planes[c] for c in order], ["bcl"] * ncyc, td.filt, widx[:500])
    assert np.array_equal(codes, wcodes) and np.array_equal(pf, wpf)
    with pytest.raises(ValueError):
        eng.tile_put_bcl(0, 0, td.planes[0])                  # a mapped slot takes no copies
    with pytest.raises(ValueError):
        eng.tile_map_host(1, n, np.zeros((ncyc, n), np.uint8))   # pageable memory is refused
    # CBCL, early cycles with every well, later ones PF wells only
    pfmask = (td.filt & 1).astype(bool)
    stride = (n + 1) // 2
    pin2 = PinnedArray((ncyc, stride))
    planes, kinds, n_block = [], [], []
    for c in range(ncyc):
        nib = synth.bcl_to_nibbles(td.planes[c])
        excl = c >= split
        if excl:
            nib = nib[pfmask]
        packed = synth.pack_nibbles(nib)
        pin2.array[c, : packed.size] = packed
        planes.append(packed)
        kinds.append("cbcl_excl" if excl else "cbcl")
        n_block.append(nib.size)
    wpt, wc = CP.count_tile(planes, kinds, td.filt, centres, woffs, widx, 5, 2, False)
    pinf = PinnedArray((n,))
    pinf.array[:] = td.filt
    for mapped_filter in (False, True):
        eng.tile_map_host(0, n, pin2.array, kinds=[CP.KIND[k] for k in kinds], n_block=n_block,
                          pinned_filter=pinf.array if mapped_filter else None)
        if not mapped_filter:
            eng.tile_put_filter(0, td.filt)
        for mode in (0, 1):
            pt, cnt = eng.count(0, 1, order, 2, False, mode=mode)
            assert np.array_equal(pt[0], wpt) and np.array_equal(cnt[0], wc)
        off, passing = eng.filter_offsets(0)
        woff, wpass = CP.filter_offsets(td.filt)
        assert np.array_equal(off, woff) and passing == wpass
    with pytest.raises(AssertionError):
        eng.tile_map_host(0, n, pin2.array, kinds=[CP.KIND[k] for k in kinds], n_block=[n + 1] * ncyc)
    pin.free()
    pin2.free()
    pinf.free()


def test_caller_owned_memory_can_be_mapped_after_registration(eng, oracle):
    """wd_host_register: an ordinary numpy block [tile][plane][stride] becomes usable by wd_tile_map_host
    (zero-copy staging for callers that do not allocate through wd_host_alloc); unregistered memory is refused."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import RegisteredArray
    X, Y, td, centres = _synthetic_tile(21, 50000, 250, 20, 300, dup_rate=0.3, shift_share=0.3)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    stride = (td.n_wells + 255) // 256 * 256
    block = np.zeros((2, td.n_cycles, stride), np.uint8)
    block[:, :, :td.n_wells] = td.planes
    filt = np.ascontiguousarray(td.filt)
    order = list(range(td.n_cycles))
    with pytest.raises(ValueError):
        eng.tile_map_host(0, td.n_wells, block[0], pinned_filter=filt)          # pageable: refused
    reg, regf = RegisteredArray(block), RegisteredArray(filt)
    try:
        for k in range(2):
            eng.tile_map_host(k, td.n_wells, block[k], pinned_filter=filt)
        pt, cnt = eng.count(0, 2, order, 2, False, mode=0)
        wpt, wc = CP.count_tile([td.planes[c] for c in order], ["bcl"] * len(order), td.filt, centres, offs, idx, 5, 2, False)
        for k in range(2):
            assert np.array_equal(pt[k], wpt) and np.array_equal(cnt[k], wc)
    finally:
        eng.sync()
        reg.release()
        regf.release()


@pytest.mark.parametrize("seed", range(24))
def test_random_configurations_vs_oracle(eng, oracle, seed):
    """Seeded sweep over what the fused kernel's paths depend on: compared length 1..300 (1 to 8 words, centre
    reads that straddle word boundaries), plane order with gaps and repeats, -e 0..7, Hamming or Levenshtein,
    1..5 levels, BCL or CBCL (both block kinds), planes staged in HBM or left in page-locked host memory (head
    planes + sector pulls), one or two tiles per launch, round sizes from wd_set_tuning.  Counters, per-target rows
    and the duplicate-pair log (fused and two-pass) against the C oracle."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    from well_duplicates_b200.engine import PinnedArray
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(4000, 24000))
    row_len = int(rng.integers(60, 200))
    ncyc = int(rng.choice([6, 17, 33, 50, 64, 65, 90, 130]))
    levels = int(rng.integers(1, 6))
    e = int(rng.integers(0, 8))
    ham = bool(rng.integers(0, 2))
    cbcl = bool(rng.integers(0, 3) == 0)
    mapped = bool(rng.integers(0, 2))
    n_tiles = int(rng.integers(1, 3))
    X, Y = synth.hex_lattice(n, row_len)
    centres = rng.choice(n, size=int(rng.integers(20, 200)), replace=False).astype(np.uint32)
    offs, idx = CP.rings_csr(X, Y, centres, levels)
    if seed % 3 == 0:
        # a list with more levels than prepare_cluster_indexes writes (the library takes up to 15): every target's wells
        # regrouped into 6..9 consecutive, non-empty levels
        levels = int(rng.integers(6, 10))
        offs5, idx = CP.rings_csr(X, Y, centres, 5)
        cuts = [0]
        for t in range(centres.size):
            a, b = int(offs5[5 * t]), int(offs5[5 * t + 5])
            inner = np.sort(rng.choice(np.arange(a + 1, b), size=levels - 1, replace=False))
            cuts += inner.tolist() + [b]
        offs = np.array(cuts, np.uint32)
    eng.load_targets(centres, offs, idx, levels)
    # compared positions: a few ranges of planes, possibly overlapping and out of order (--cycles a-b,c-d)
    order = []
    for _ in range(int(rng.integers(1, 4))):
        a = int(rng.integers(0, ncyc))
        order += list(range(a, int(rng.integers(a + 1, ncyc + 1))))
    order = order[:300]
    split = int(rng.integers(0, ncyc + 1))
    tiles, pins = [], []
    for k in range(n_tiles):
        td = synth.make_tile(rng, n, ncyc, row_len, pf_rate=float(rng.uniform(0.3, 1.0)), dup_rate=float(rng.uniform(0.05, 0.5)),
                             shift_share=float(rng.uniform(0, 0.6)), nocall_rate=float(rng.uniform(0, 0.05)))
        pfmask = (td.filt & 1).astype(bool)
        planes, kinds, nb = [], [], []
        for c in range(ncyc):
            if not cbcl:
                planes.append(td.planes[c]); kinds.append("bcl"); nb.append(n)
                continue
            nib = synth.bcl_to_nibbles(td.planes[c])
            if c >= split:
                nib = nib[pfmask]
            planes.append(synth.pack_nibbles(nib)); kinds.append("cbcl_excl" if c >= split else "cbcl"); nb.append(nib.size)
        tiles.append((planes, kinds, nb, td.filt))
    stride = (n + 255) // 256 * 256
    if mapped:
        block = PinnedArray((n_tiles, ncyc, stride))
        block.array[:] = 0
        fpin = PinnedArray((n_tiles, stride))
        pins += [block, fpin]
    for k, (planes, kinds, nb, filt) in enumerate(tiles):
        if mapped:
            for c in range(ncyc):
                block.array[k, c, :planes[c].size] = planes[c]
            fpin.array[k, :n] = filt
            eng.tile_map_host(k, n, block.array[k], kinds=[CP.KIND[x] for x in kinds], n_block=nb, pinned_filter=fpin.array[k, :n])
        else:
            eng.tile_begin(k, n, ncyc)
            eng.tile_put_filter(k, filt)
            for c in range(ncyc):
                if kinds[c] == "bcl":
                    eng.tile_put_bcl(k, c, planes[c])
                else:
                    eng.tile_put_cbcl(k, c, planes[c], nb[c], kinds[c] == "cbcl_excl")
    if rng.integers(0, 2):
        eng.set_tuning(step0=int(rng.integers(1, 9)), step1=int(rng.integers(1, 9)), head_planes=int(rng.integers(0, 4)),
                       centre_chunk=int(rng.choice([0, 8, 16, 32])), visit_order=int(rng.integers(0, 2)),
                       targets_per_cta=int(rng.choice([0, 8, 24, 40, 64, 256])), ctas_per_sm=int(rng.choice([0, 0, 1, 3])))
    try:
        want = [CP.count_tile([pl[c] for c in order], [kd[c] for c in order], filt, centres, offs, idx, levels, e, ham)
                for pl, kd, nb, filt in tiles]
        logs = {}
        for mode in (0, 2, 1):
            pt, cnt = eng.count(0, n_tiles, order, e, ham, mode=mode)
            for k in range(n_tiles):
                assert np.array_equal(pt[k], want[k][0]) and np.array_equal(cnt[k], want[k][1]), (seed, mode, k)
            if mode:
                logs[mode] = eng.dup_pairs(with_seqs=True)
        assert np.array_equal(logs[1][0], logs[2][0]) and np.array_equal(logs[1][1], logs[2][1])
        assert len(logs[2][0]) == int(sum(w[1][2::5].sum() for w in want))
        if len(logs[2][0]):
            rows, codes = logs[2]
            k = int(rows[0, 0])
            pl, kd, nb, filt = tiles[k]
            wcodes, _ = CP.get_codes([pl[c] for c in order], [kd[c] for c in order], filt, rows[rows[:, 0] == k][:, 2].astype(np.int64))
            assert np.array_equal(codes[rows[:, 0] == k][:, 1], wcodes)
    finally:
        eng.set_tuning()
        eng.sync()
        for pin in pins:
            pin.free()


def test_dup_pair_log_rows(eng, oracle):
    """Rows behind the stderr log: (tile, target, well, distance) in reference order."""
    R, CP = oracle
    case = [c for c in MAN["count"] if c["name"] == "long75_e3"][0]
    o = parse_count_args(case["args"])
    order, tiles, (centres, offs, idx), lane = _load_case(eng, R, case, o)
    eng.count(0, len(tiles), order, o["edit"], o["hamming"], mode=1)
    rows = eng.dup_pairs()
    # the fused kernel logs the same pairs with the same distances, and hands out both sequences
    eng.count(0, len(tiles), order, o["edit"], o["hamming"], mode=2)
    eng.get_seqs(len(tiles) - 1, [3, 1], order[2:7])       # another gather in between must not disturb the log's sequences
    rows2, codes = eng.dup_pairs(with_seqs=True)
    assert np.array_equal(rows, rows2)
    log = []
    run = os.path.join(GOLDEN, case["run"])
    targets = R.parse_target_file(os.path.join(GOLDEN, case["targets"]), levels=o["levels"] + 1, limit=o["limit"])
    want = []
    for k, t in enumerate(tiles):
        seq_objs = [R.get_seqs_run(run, lane, t, R.all_indices(targets), s, e) for s, e in o["ranges"]]
        log = []
        R.count_tile(targets, seq_objs, o["levels"], o["edit"], o["hamming"], log)
        want += [(k, c, w, d) for c, _, w, _, d in log]
    got = [(int(r[0]), int(centres[r[1]]), int(r[2]), int(r[3])) for r in rows]
    assert got == want and len(got) > 0
    from well_duplicates_b200.reader import codes_to_strings
    seqs = codes_to_strings(codes.reshape(-1, len(order)))
    k = 0
    for ti, t in enumerate(tiles):
        seq_objs = [R.get_seqs_run(run, lane, t, R.all_indices(targets), s, e) for s, e in o["ranges"]]
        log = []
        R.count_tile(targets, seq_objs, o["levels"], o["edit"], o["hamming"], log)
        for c, cseq, w, wseq, d in log:
            assert (seqs[2 * k], seqs[2 * k + 1]) == (cseq, wseq)
            k += 1
    assert k == len(rows)


@pytest.mark.parametrize("e,ham", [(2, False), (1, False), (0, False), (3, True), (5, False), (60, False), (70, True)])
def test_fused_log_equals_two_pass_log(eng, oracle, e, ham):
    """Distances logged by the fused kernel (read off its prefix programme) against the two-pass
    kernel's exact_distance, incl. e >= len where every pair is a duplicate and must still be measured."""
    from well_duplicates_b200 import synth
    X, Y, td, centres = _synthetic_tile(11, 60000, 300, 50, 400, dup_rate=0.4, shift_share=0.5, nocall_rate=0.01)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    eng.tile_begin(0, td.n_wells, td.n_cycles)
    eng.tile_put_filter(0, td.filt)
    for c in range(td.n_cycles):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(td.n_cycles))
    pt1, c1 = eng.count(0, 1, order, e, ham, mode=1)
    rows1 = eng.dup_pairs()
    pt2, c2 = eng.count(0, 1, order, e, ham, mode=2)
    rows2 = eng.dup_pairs()
    assert np.array_equal(pt1, pt2) and np.array_equal(c1, c2)
    assert np.array_equal(rows1, rows2) and len(rows1) == int(c1[0, 2::5].sum()) > 0


def test_dup_log_grows_when_every_ring_well_is_a_duplicate(eng, oracle):
    """A tile of identical reads: every ring well of every valid target is a duplicate (amplicon-like
    libraries, large -e).  The device log starts smaller than that; wd_dup_pairs grows it and repeats the
    count instead of failing -- the reference logs every pair (count_well_duplicates.py:258-262)."""
    from well_duplicates_b200 import synth
    n, row_len, ncyc, t = 250000, 500, 20, 2400
    rng = np.random.default_rng(5)
    X, Y = synth.hex_lattice(n, row_len)
    centres = rng.choice(n, size=t, replace=False).astype(np.uint32)
    planes = np.repeat(rng.integers(1, 4, size=(ncyc, 1), dtype=np.uint8) | 0x40, n, axis=1)
    filt = np.ones(n, np.uint8)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    eng.tile_begin(0, n, ncyc)
    eng.tile_put_filter(0, filt)
    for c in range(ncyc):
        eng.tile_put_bcl(0, c, planes[c])
    for mode in (2, 1):
        pt, cnt = eng.count(0, 1, list(range(ncyc)), 2, False, mode=mode)
        rows = eng.dup_pairs()
        assert len(rows) == idx.size > 65536 + (idx.size + t) // 8       # more than the log's first size
        assert np.array_equal(rows[:, 2], idx) and not rows[:, 3].any() and not rows[:, 0].any()
        assert np.array_equal(rows[:, 1], np.repeat(np.arange(t), np.diff(offs[::5].astype(np.int64))))
        assert np.array_equal(cnt[0, 1::5], cnt[0, 2::5])                # Wells == Dups at every level


def test_empty_ring_is_reported_only_for_valid_centres(eng):
    """count_well_duplicates.py:249 asserts a ring holds wells when it gets to it -- i.e. for a target whose
    centre passes the filter of the tile being counted; a list with an empty ring on a never-valid centre runs."""
    eng.load_targets([5, 9], [0, 2, 2, 3, 4], [1, 2, 8, 10], 2)          # target 0: ring 2 empty
    eng.tile_begin(0, 100, 1)
    filt = np.ones(100, np.uint8)
    filt[5] = 0
    eng.tile_put_filter(0, filt)
    eng.tile_put_bcl(0, 0, np.full(100, 5, np.uint8))
    for mode in (0, 1, 2):
        pt, cnt = eng.count(0, 1, [0], 2, False, mode=mode)
        assert cnt[0, 0] == 1 and pt[0, 0, 0] == 0 and pt[0, 1].tolist() == [1, 1, 1, 1, 1]
    filt[5] = 1
    eng.tile_put_filter(0, filt)
    for mode in (0, 1, 2):
        with pytest.raises(AssertionError):
            eng.count(0, 1, [0], 2, False, mode=mode)


def test_publish_counters_rows(eng, oracle):
    """K7 (wd_publish_counters): every tile row lands where the map says, lane rows hold the sums of their
    tiles, every other row stays zero; the single-rank all-reduce leaves the buffer as it is."""
    R, CP = oracle
    case = [c for c in MAN["count"] if c["name"] == "lev_default"][0]
    o = parse_count_args(case["args"])
    order, tiles, _, lane = _load_case(eng, R, case, o)
    _, cnt = eng.count(0, len(tiles), order, o["edit"], o["hamming"], mode=0, per_target=False)
    n_rows = 11
    tile_row = np.array([7, 2][:len(tiles)], np.int32)
    lane_row = np.array([9, 10][:len(tiles)], np.int32)
    ptr, n = eng.publish_counters(tile_row, lane_row, n_rows)
    assert n == n_rows * cnt.shape[1]
    eng.allreduce_published()
    buf = eng.published_fetch(n).reshape(n_rows, -1)
    want = np.zeros_like(buf)
    for k in range(len(tiles)):
        want[tile_row[k]] = cnt[k]
        want[lane_row[k]] += cnt[k]
    assert np.array_equal(buf, want) and buf.any()
    # both tiles into one lane row
    ptr, n = eng.publish_counters(tile_row, np.full(len(tiles), 4, np.int32), n_rows)
    buf = eng.published_fetch(n).reshape(n_rows, -1)
    assert np.array_equal(buf[4], cnt.sum(axis=0))


@pytest.mark.parametrize("chunk", [0, 8])
def test_sector_trace_equals_a_host_replay(eng, oracle, chunk):
    """wd_count_trace_sectors (bench.py's roofline numerator): with the schedule pinned to one cycle per round,
    the sectors the fused kernel reads at position p are those of the centres (as far as the programme looks
    ahead) and of the ring wells whose first p symbols have not yet proved dist > e -- replayed with the oracle."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    X, Y, td, centres = _synthetic_tile(3, 40000, 250, 24, 200, dup_rate=0.3, shift_share=0.4)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    eng.tile_begin(0, td.n_wells, td.n_cycles)
    eng.tile_put_filter(0, td.filt)
    for c in range(td.n_cycles):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(td.n_cycles))
    e, k, L = 2, 1, td.n_cycles
    eng.set_tuning(step0=1, step1=1, centre_chunk=chunk)
    try:
        sectors, lines = eng.trace_sectors(0, 1, order, e, False)
    finally:
        eng.set_tuning()
    codes, _ = CP.get_codes([td.planes[c] for c in order], ["bcl"] * L, td.filt, np.arange(td.n_wells, dtype=np.int64))
    need = [set() for _ in range(L)]
    for t, c in enumerate(centres):
        if not td.filt[c] & 1:
            continue
        ring = idx[offs[5 * t]:offs[5 * t + 5]]
        deepest = 0
        for w in ring:
            p = 0                                    # symbols of w read before its prefix proves dist > e
            while p < L:
                p += 1
                if CP.prefix_band_min(codes[c], codes[w], p, k) > e:
                    break
            for q in range(p):
                need[q].add(int(w) >> 5)
            deepest = max(deepest, p)
        # the centre is read as far as the deepest round looked ahead (p + k) -- exactly, or in chunks of 8 cycles
        known = min(L, deepest + k)
        if chunk:
            known = min(L, (known + chunk - 1) // chunk * chunk)
        for q in range(known):
            need[q].add(int(c) >> 5)
    assert sectors[0].tolist() == [len(s) for s in need]
    assert (lines[0] <= sectors[0]).all() and (4 * lines[0] >= sectors[0]).all()


def test_error_paths(eng):
    with pytest.raises(ValueError):
        eng.count(0, 1, [0] * 2000, 2, False)                 # longer than WD_MAX_SEQ_LEN
    with pytest.raises(ValueError):
        eng.count(50000, 1, [0], 2, False)                    # slot never begun
    eng.load_targets([5], [0, 2], [1, 200], 1)
    eng.tile_begin(0, 100, 1)
    eng.tile_put_filter(0, np.ones(100, np.uint8))
    eng.tile_put_bcl(0, 0, np.full(100, 5, np.uint8))
    with pytest.raises(IndexError):
        eng.count(0, 1, [0], 2, False)                        # target well 200 on a 100-well tile


# ------------------------------------------------------------ exhaustive mode --
def _exhaustive_manifest():
    with open(os.path.join(GOLDEN, "exhaustive", "manifest.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("case", _exhaustive_manifest(), ids=lambda c: c["name"])
def test_exhaustive_matches_reference_report(eng, oracle, case, tmp_path):
    """wd_count_exhaustive (every well a target, neighbourhoods straight from the
    stage-1 grid) prints what the unmodified reference printed for
    prepare_cluster_indexes.py -n <all wells> + count_well_duplicates.py."""
    R, CP = oracle
    from well_duplicates_b200 import report
    from well_duplicates_b200.reader import BCLReader
    o = parse_count_args(case["args"])
    _, xy = R.read_locs(locs_path(case["locs"], tmp_path))
    eng.load_locs(xy)
    rd = BCLReader(os.path.join(GOLDEN, "run_bcl"), engine=eng)
    wanted = [c for s, e in o["ranges"] for c in range(s, e)]
    rows = []
    for k, t in enumerate(case["tiles"]):
        plane_of = rd.get_tile(case["lane"], t).stage(k, wanted)
        rows.append(eng.count_exhaustive(k, [plane_of[c] for c in wanted], o["levels"], o["edit"], o["hamming"]))
    with open(os.path.join(GOLDEN, "exhaustive", case["name"] + ".stdout")) as fh:
        want = fh.read()
    assert report.format_report(case["lane"], xy.shape[0], case["tiles"], rows, o["levels"], verbose=True) == want


@pytest.mark.parametrize("case", _exhaustive_manifest()[:2], ids=lambda c: c["name"])
def test_every_well_through_a_binary_target_list(case, tmp_path, capsys):
    """The reference's own route to the same report -- prepare_cluster_indexes.py -n <all wells> -s 1, then
    count_well_duplicates.py -f <that list> -n <all wells> -- with the list written in binary form (SURVEY 8 f3:
    the text form of a full tile is 3 GB) and counted by the sampled-mode kernels: same stdout as the reference."""
    from well_duplicates_b200 import count_cli, prepare_cli
    binary = str(tmp_path / "all_wells.bin")
    prepare_cli.main(["-f", locs_path(case["locs"], tmp_path), "-n", "3072", "-s", "1", "--binary", binary])
    capsys.readouterr()
    count_cli.main(["-f", binary, "-n", "3072", "-r", os.path.join(GOLDEN, "run_bcl"), "-s", "hiseq_x", "-i", case["lane"],
                    "-t", ",".join(case["tiles"]), "-q"] + case["args"])
    with open(os.path.join(GOLDEN, "exhaustive", case["name"] + ".stdout")) as fh:
        assert capsys.readouterr().out == fh.read()


@pytest.mark.parametrize("case", _exhaustive_manifest(), ids=lambda c: c["name"])
def test_exhaustive_cli_matches_reference_report(case, tmp_path):
    """count_well_duplicates.py --exhaustive-locs S_LOCS (an extension: no target file) prints what the
    unmodified reference printed for a target file holding every well."""
    from well_duplicates_b200 import count_cli
    argv = ["--exhaustive-locs", locs_path(case["locs"], tmp_path), "-r", os.path.join(GOLDEN, "run_bcl"), "-s", "hiseq_x",
            "-i", case["lane"], "-t", ",".join(case["tiles"]), "-q"] + case["args"]
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        count_cli.main(argv)
    with open(os.path.join(GOLDEN, "exhaustive", case["name"] + ".stdout")) as fh:
        assert out.getvalue() == fh.read()


@pytest.mark.parametrize("e,ham,levels", [(2, False, 5), (2, True, 5), (3, False, 2)])
def test_exhaustive_medium_tile_vs_oracle(eng, oracle, e, ham, levels):
    """A cropped tile (150 rows x 300 wells, 50 cycles), every well a target,
    against the C oracle; also equal to the sampled path fed with all wells."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(31)
    n, row_len, ncyc = 45000, 300, 50
    X, Y = synth.hex_lattice(n, row_len)
    td = synth.make_tile(rng, n, ncyc, row_len, dup_rate=0.1, shift_share=0.4, nocall_rate=0.003)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    eng.tile_begin(0, n, ncyc)
    eng.tile_put_filter(0, td.filt)
    for c in range(ncyc):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(ncyc))
    got = eng.count_exhaustive(0, order, levels, e, ham)
    want = CP.count_exhaustive(X, Y, [td.planes[c] for c in order], ["bcl"] * ncyc, td.filt, levels, e, ham)
    assert np.array_equal(got, want)
    centres = np.arange(n, dtype=np.uint32)
    offs, idx = eng.ring_query(centres, levels)
    eng.load_targets(centres, offs, idx, levels)
    _, cnt = eng.count(0, 1, order, e, ham, mode=0, per_target=False)
    assert np.array_equal(cnt[0], got)
    assert got[2::5].sum() > 100


def test_exhaustive_index_window_edge(eng, oracle):
    """Rows of 4000 wells: the wells five rows away sit at index distance
    20000 +- 2, exactly where the reference's scan window [c - 20000, c + 20001]
    ends (prepare_cluster_indexes.py:52-67).  The window is not symmetric, so a
    duplicate pair may count for one of its wells and not for the other."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(77)
    row_len, rows, ncyc = 4000, 12, 20
    n = row_len * rows
    X, Y = synth.hex_lattice(n, row_len)
    td = synth.make_tile(rng, n, ncyc, row_len, dup_rate=0.0)
    # plant copies five rows up and down, at every index offset around the window edge
    planes = td.planes
    for k, off in enumerate([19998, 19999, 20000, 20001, 20002] * 40):
        a = int(rng.integers(0, n - 20010))
        planes[:, a + off] = planes[:, a]
    td.filt[:] = 1
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    eng.tile_begin(0, n, ncyc)
    eng.tile_put_filter(0, td.filt)
    for c in range(ncyc):
        eng.tile_put_bcl(0, c, planes[c])
    order = list(range(ncyc))
    for ham in (False, True):
        got = eng.count_exhaustive(0, order, 5, 2, ham)
        want = CP.count_exhaustive(X, Y, [planes[c] for c in order], ["bcl"] * ncyc, td.filt, 5, 2, ham)
        assert np.array_equal(got, want)
    assert got[1 + 5 * 4 + 1] > 50            # ring-5 duplicates were found at all


def test_exhaustive_low_complexity_reads(eng, oracle):
    """Blocks of no-call wells (all N) beside poly-A wells: in the 32-symbol
    pre-test an N reads as A, so these pairs all reach the exact compare, which
    must tell them apart (N is an ordinary symbol, count_well_duplicates.py:251-252)."""
    R, CP = oracle
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(78)
    n, row_len, ncyc = 24000, 200, 40
    X, Y = synth.hex_lattice(n, row_len)
    td = synth.make_tile(rng, n, ncyc, row_len, dup_rate=0.05, shift_share=0.5, nocall_rate=0.01)
    for r0 in (10, 40, 41, 42, 90):
        td.planes[:, r0 * row_len + 20:r0 * row_len + 60] = 0            # all N
        td.planes[:, r0 * row_len + 60:r0 * row_len + 90] = 0x5c         # poly-A, quality 23
        td.planes[:, r0 * row_len + 90:r0 * row_len + 110] = 0x5f        # poly-T
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    eng.tile_begin(0, n, ncyc)
    eng.tile_put_filter(0, td.filt)
    for c in range(ncyc):
        eng.tile_put_bcl(0, c, td.planes[c])
    order = list(range(ncyc))
    for e, ham in ((2, False), (0, True), (5, False), (1, False)):
        got = eng.count_exhaustive(0, order, 5, e, ham)
        want = CP.count_exhaustive(X, Y, [td.planes[c] for c in order], ["bcl"] * ncyc, td.filt, 5, e, ham)
        assert np.array_equal(got, want), (e, ham)
    for e in (-1, 40, 100):                      # no pair / every pair is a duplicate: no sequence is compared
        got = eng.count_exhaustive(0, order, 3, e, False)
        want = CP.count_exhaustive(X, Y, [td.planes[c] for c in order], ["bcl"] * ncyc, td.filt, 3, e, False)
        assert np.array_equal(got, want), e


def test_exhaustive_errors(eng, oracle):
    R, CP = oracle
    from well_duplicates_b200 import synth
    # three wells far apart: every ring is empty -> the reference's RuntimeError
    eng.load_locs(fx.sparse3())
    eng.tile_begin(0, 3, 2)
    eng.tile_put_filter(0, np.ones(3, np.uint8))
    for c in range(2):
        eng.tile_put_bcl(0, c, np.full(3, 5, np.uint8))
    with pytest.raises(RuntimeError, match="Got no wells"):
        eng.count_exhaustive(0, [0, 1])
    # .locs and tile disagree on the number of wells
    eng.load_locs(fx.hex_tiny())
    with pytest.raises(AssertionError):
        eng.count_exhaustive(0, [0, 1])


# ------------------------------------------------------- full-size properties --
@pytest.fixture(scope="module")
def full_tile(eng):
    """One HiSeq 4000 tile at BASELINE size: 4 309 650 wells, 50 cycles, 2500 targets."""
    from well_duplicates_b200 import synth
    n, row_len, ncyc = synth.HISEQ4000_WELLS, synth.HISEQ4000_ROW_LEN, 50
    rng = np.random.default_rng(20261018)
    X, Y = synth.hex_lattice(n, row_len)
    td = synth.make_tile(rng, n, ncyc, row_len)
    import random
    random.seed(13)
    centres = np.array(random.sample(range(n), 2500), dtype=np.uint32)
    eng.load_locs(synth.xy_to_locs_floats(X, Y))
    offs, idx = eng.ring_query(centres, 5)
    eng.load_targets(centres, offs, idx, 5)
    eng.tile_begin(0, n, ncyc)
    eng.tile_put_filter(0, td.filt)
    for c in range(ncyc):
        eng.tile_put_bcl(0, c, td.planes[c])
    return X, Y, td, centres, offs, idx


def test_full_tile_stage1_vs_oracle(full_tile, oracle):
    R, CP = oracle
    X, Y, td, centres, offs, idx = full_tile
    woffs, widx = CP.rings_csr(X, Y, centres[:400])
    assert np.array_equal(offs[: woffs.size], woffs) and np.array_equal(idx[: widx.size], widx)
    lens = np.diff(offs.astype(np.int64))
    assert lens.min() >= 1
    # ascending inside every ring
    seg = np.repeat(np.arange(lens.size), lens)
    d = np.diff(idx.astype(np.int64))
    assert np.all(d[seg[1:] == seg[:-1]] > 0)


@pytest.mark.parametrize("ham", [False, True])
def test_full_tile_count_vs_oracle_and_properties(eng, full_tile, oracle, ham):
    R, CP = oracle
    X, Y, td, centres, offs, idx = full_tile
    order = list(range(50))
    pt0, c0 = eng.count(0, 1, order, 2, ham, mode=0)
    pt1, c1 = eng.count(0, 1, order, 2, ham, mode=1)
    assert np.array_equal(pt0, pt1) and np.array_equal(c0, c1)                       # two kernels, one answer
    pt2, c2 = eng.count(0, 1, order, 2, ham, mode=0)
    assert np.array_equal(pt0, pt2) and np.array_equal(c0, c2)                       # idempotent
    wpt, wc = CP.count_tile([td.planes[c] for c in order], ["bcl"] * 50, td.filt, centres, offs, idx, 5, 2, ham)
    assert np.array_equal(pt0[0], wpt) and np.array_equal(c0[0], wc)                 # the oracle
    valid = pt0[0][:, 0] == 1
    assert np.array_equal(valid, (td.filt[centres] & 1) == 1)                        # centre PF rule
    assert c0[0][0] == valid.sum()
    lens = np.diff(offs.astype(np.int64)).reshape(-1, 5)
    assert np.array_equal(c0[0][1::5], lens[valid].sum(axis=0))                      # Wells = ring sizes of valid targets
    assert np.array_equal(c0[0][2::5], pt0[0][:, 1::2].sum(axis=0))                  # Dups = sum of tallies
    assert c0[0][4 + 5 * 4] == c0[0][5] == (pt0[0][:, 1::2].sum(axis=1) > 0).sum()   # AccO[5] == AccI[1] == any hit


def test_full_tile_monotone_in_e(eng, full_tile):
    order = list(range(50))
    prev = None
    for e in (0, 1, 2, 3, 5):
        pt, _ = eng.count(0, 1, order, e, False, mode=0)
        ph, _ = eng.count(0, 1, order, e, True, mode=0)
        assert np.all(pt[0][:, 1::2] >= ph[0][:, 1::2])                              # Lev <= Ham
        if prev is not None:
            assert np.all(pt[0][:, 1::2] >= prev)
        prev = pt[0][:, 1::2]


def test_full_tile_exhaustive_properties(eng, full_tile):
    """Exhaustive mode at BASELINE size (every one of the 4 309 650 wells a target):
    properties that do not need an oracle run of that size."""
    X, Y, td, centres, offs, idx = full_tile
    order = list(range(50))
    lev = eng.count_exhaustive(0, order, 5, 2, False)
    assert np.array_equal(lev, eng.count_exhaustive(0, order, 5, 2, False))          # idempotent (tallies are reset)
    ham = eng.count_exhaustive(0, order, 5, 2, True)
    n_pf = int((td.filt & 1).sum())
    assert lev[0] == ham[0] == n_pf                                                  # centre PF rule
    assert np.array_equal(lev[1::5], ham[1::5])                                      # Wells are geometry only
    assert np.all(lev[2::5] >= ham[2::5]) and lev[2::5].sum() > ham[2::5].sum()      # Lev <= Ham, planted shifts differ
    assert lev[4 + 5 * 4] == lev[5]                                                  # AccO[5] == AccI[1] == any hit
    assert np.all(np.diff(lev[4::5]) >= 0) and np.all(np.diff(lev[5::5]) <= 0)       # AccO grows outwards, AccI inwards
    # interior wells of the lattice have 6 k wells in ring k: the mean ring size sits just below that
    assert np.all(lev[1::5] <= 6 * np.arange(1, 6) * n_pf) and np.all(lev[1::5] >= 5.9 * np.arange(1, 6) * n_pf)
    # fewer levels: the same inner rings
    l3 = eng.count_exhaustive(0, order, 3, 2, False)
    assert l3[0] == lev[0]
    for k in (1, 2, 3, 4):                                                           # Wells, Dups, Hit, AccO of rings 1..3
        assert np.array_equal(l3[k::5], lev[k::5][:3])
    # the 2500 sampled targets are a subset: their duplicates cannot exceed the exhaustive totals
    _, c = eng.count(0, 1, order, 2, False, mode=0, per_target=False)
    assert np.all(c[0][2::5] <= lev[2::5]) and np.all(c[0][3::5] <= lev[3::5])
    # no sequence decides: e >= len makes every ring well a duplicate, e < 0 none
    every = eng.count_exhaustive(0, order, 5, 50, False)
    assert np.array_equal(every[2::5], every[1::5]) and np.all(every[3::5] == n_pf) and np.all(every[4::5] == n_pf)
    none = eng.count_exhaustive(0, order, 5, -1, False)
    assert none[0] == n_pf and not none[2::5].any() and not none[3::5].any() and np.array_equal(none[1::5], lev[1::5])
    # monotone in e
    e1 = eng.count_exhaustive(0, order, 5, 1, False)
    e3 = eng.count_exhaustive(0, order, 5, 3, False)
    assert np.all(e1[2::5] <= lev[2::5]) and np.all(lev[2::5] <= e3[2::5])


# ------------------------------------------------------------------ flowcell driver --
@pytest.mark.parametrize("name", ["two_lanes", "lev_default", "cbcl_default", "summary"])
def test_flowcell_driver_single_rank(name):
    """One process, one GPU: the whole-run driver prints what the reference prints
    lane after lane (the multi-rank exchange is covered on CPU by test_dist_gloo
    and on 2 GPUs by tests/multi_gpu_check.sh)."""
    from well_duplicates_b200 import flowcell
    case = [c for c in MAN["count"] if c["name"] == name][0]
    with open(os.path.join(GOLDEN, "count", name + ".stdout")) as fh:
        want = fh.read()
    argv = ["-f", os.path.join(GOLDEN, case["targets"]), "-r", os.path.join(GOLDEN, case["run"])] + case["args"]
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        flowcell.main(argv + ["-q"])
    assert out.getvalue() == want
    # without -q a rank logs its tiles as the reference does; on one rank that is the reference's stderr
    with open(os.path.join(GOLDEN, "count", name + ".stderr")) as fh:
        want_err = fh.read()
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        flowcell.main(argv)
    assert out.getvalue() == want and err.getvalue() == want_err


# ------------------------------------------------------- whole-run workflow --
def test_workflow_whole_run(tmp_path):
    """python -m well_duplicates_b200.workflow RUN WORKDIR: the per-lane files are
    what the count command prints for the flags Snakefile.count_dups:153-160
    derives from RunInfo.xml, the summary is their ``tail`` (:146-151)."""
    import shutil
    import subprocess
    from well_duplicates_b200 import count_cli, workflow
    run = tmp_path / "run"
    run.mkdir()
    os.symlink(os.path.join(GOLDEN, "run_bcl", "Data"), run / "Data")
    (run / "RunInfo.xml").write_text(
        '<RunInfo><Run><Reads><Read Number="1" NumCycles="14"/></Reads><FlowcellLayout><TileSet><Tiles>'
        '<Tile>1_1101</Tile><Tile>2_1101</Tile></Tiles></TileSet>'
        '</FlowcellLayout></Run></RunInfo>')
    work = tmp_path / "work"
    work.mkdir()
    shutil.copy(os.path.join(GOLDEN, "locs", "hex_small_n40_s13.list"), work / "40clusters.list")
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        workflow.main([str(run), str(work), "-n", "40", "--read-length", "13"])
    names = ["40targets_lane1.txt", "40targets_lane2.txt"]
    for lane, name in zip("12", names):
        out = io.StringIO()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
            count_cli.main(["-f", str(work / "40clusters.list"), "-n", "40", "-s", "1101", "-r", str(run), "-i", lane,
                            "-l", "5", "--cycles", "0-13"])
        assert (work / name).read_text() == out.getvalue()
        assert "1101" in out.getvalue()
    want = subprocess.run(["tail", "-n", "6"] + names, capture_output=True, text=True, cwd=work).stdout
    assert (work / "40targets_all_lanes.txt").read_text() == want


def test_staging_pipeline_zero_copy_and_copied_equal_oracle(eng, oracle, tmp_path):
    """staging.lane_batches: files -> page-locked block (native inflate threads) -> kernels.
    Planes read in place across PCIe and planes copied from the block give the oracle's
    counters, for BCL tiles of two sizes and a CBCL lane with both block kinds, whatever the
    batch size; get_seqs through the same path equals the oracle's decode."""
    R, CP = oracle
    from well_duplicates_b200 import staging, synth
    from well_duplicates_b200.reader import BCLReader
    rng = np.random.default_rng(21)
    run = str(tmp_path / "run")
    row_len, ncyc = 120, 24
    sizes = {1101: 20011, 1102: 20011, 1103: 26000, 1104: 26000, 1105: 26000}
    data = {1: {}, 2: {}}
    for tile, n in sizes.items():
        data[1][tile] = synth.make_tile(rng, n, ncyc, row_len, pf_rate=0.7, dup_rate=0.25, shift_share=0.3, nocall_rate=0.01)
        synth.write_bcl_tile(run, 1, tile, data[1][tile], compresslevel=int(rng.integers(1, 9)))
    for tile in (1101, 2101, 1102):
        data[2][tile] = synth.make_tile(rng, 20011, ncyc, row_len, pf_rate=0.6, dup_rate=0.25, shift_share=0.3)
    synth.write_cbcl_lane(run, 2, data[2], excluded_from_cycle=9)
    X, Y = synth.hex_lattice(20011, row_len)
    centres = rng.choice(20011, size=500, replace=False).astype(np.uint32)
    offs, idx = CP.rings_csr(X, Y, centres)
    eng.load_targets(centres, offs, idx, 5)
    wanted = list(range(3, 21)) + [1, 2]
    rd = BCLReader(run, engine=eng)
    st = staging.Stager(threads=4, cbcl_cache=rd._cbcl_cache)
    pfm = {t: (td.filt & 1).astype(bool) for t, td in data[2].items()}

    def oracle_rows(lane, tile):
        td = data[lane][tile]
        if lane == 1:
            planes, kinds = [td.planes[c] for c in wanted], ["bcl"] * len(wanted)
        else:
            planes, kinds = [], []
            for c in wanted:
                nib = synth.bcl_to_nibbles(td.planes[c])
                planes.append(synth.pack_nibbles(nib[pfm[tile]] if c >= 9 else nib))
                kinds.append("cbcl_excl" if c >= 9 else "cbcl")
        return CP.count_tile(planes, kinds, td.filt, centres, offs, idx, 5, 2, False)

    for lane, names in ((1, list(sizes)), (2, [1101, 2101, 1102])):
        want = {t: oracle_rows(lane, t) for t in names}
        for per_batch in (None, 2, 1):
            for zero_copy, mode in ((True, 0), (False, 0), (False, 1), (True, 1)):
                seen = []
                for got, batch in staging.lane_batches(st, lambda t: rd.get_tile(lane, t), names, wanted, per_batch=per_batch):
                    plane_of = st.deliver(eng, batch, first_slot=0, zero_copy=zero_copy)
                    pt, cnt = eng.count(0, len(got), [plane_of[c] for c in wanted], 2, False, mode=mode)
                    for k, t in enumerate(got):
                        assert np.array_equal(pt[k], want[t][0]) and np.array_equal(cnt[k], want[t][1]), (lane, t, per_batch, zero_copy, mode)
                    seen += got
                assert seen == names
        # the reader API on top of the same staging: a few wells (sectors pulled in place) and many (planes copied)
        t = names[-1]
        td = data[lane][t]
        for wells in (idx[:40], np.arange(0, 20011, 3)):
            got = rd.get_tile(lane, t).get_seqs([int(w) for w in wells], 2, 20)
            if lane == 1:
                planes, kinds = [td.planes[c] for c in range(2, 20)], ["bcl"] * 18
            else:
                planes = [synth.pack_nibbles(synth.bcl_to_nibbles(td.planes[c])[pfm[t]] if c >= 9 else synth.bcl_to_nibbles(td.planes[c]))
                          for c in range(2, 20)]
                kinds = ["cbcl_excl" if c >= 9 else "cbcl" for c in range(2, 20)]
            keys = sorted({int(w) for w in wells})
            wcodes, wpf = CP.get_codes(planes, kinds, td.filt, np.array(keys))
            assert [got[k][0] for k in keys] == ["".join("ACGTN"[c] for c in row) for row in wcodes]
            assert [got[k][1] for k in keys] == [bool(f) for f in wpf]
    st.close()
