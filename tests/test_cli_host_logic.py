"""Host logic of the drop-in CLI on top of the staging pipeline, without a GPU.

count_cli.main is run on every golden case with a stand-in for engine.Engine
whose kernels are the CPU oracle (test infrastructure only: the product has no
such path and refuses to run without a GPU, test_host_logic.py).  What is
checked here is everything between the files and the kernels -- tile walk,
batching, BCL/CBCL choice, native inflate into (pageable) plane blocks, the
order of the log lines -- against the stdout and stderr of the unmodified
reference.  The same cases run on the real kernels in test_gpu_parity.py."""
import contextlib
import io
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_count_output, load_manifest, run_count_case
from oracle import c_port as CP
from well_duplicates_b200 import _lib, count_cli, reader, staging

MAN = load_manifest()
KIND_NAME = {_lib.PLANE_BCL: "bcl", _lib.PLANE_CBCL: "cbcl", _lib.PLANE_CBCL_EXCL: "cbcl_excl"}


class OracleEngine:
    """engine.Engine's staging / counting surface, computed by oracle/c_port."""

    def __init__(self):
        self.slots = {}
        self.calls = []
        self.modes = []

    # -- targets -----------------------------------------------------------------------------------
    def load_targets(self, centres, level_offsets, idx, levels):
        self.centres = np.asarray(centres, np.uint32)
        self.offs = np.asarray(level_offsets, np.uint32)
        self.idx = np.asarray(idx, np.uint32)
        self.levels = levels

    # -- staging: views are kept, bytes are read when the count runs (as the kernels do) -------------
    def tile_map_host(self, slot, n_clusters, pinned_planes, kinds=None, n_block=None, pinned_filter=None):
        assert pinned_planes.ndim == 2 and pinned_planes.flags.c_contiguous and pinned_filter is not None
        self.calls.append("map")
        self.slots[slot] = dict(n=n_clusters, planes=[pinned_planes[p] for p in range(pinned_planes.shape[0])],
                                kinds=[int(k) for k in kinds], n_block=[int(b) for b in n_block], filt=pinned_filter)

    def tile_begin(self, slot, n_clusters, n_planes):
        self.calls.append("begin")
        self.slots[slot] = dict(n=n_clusters, planes=[None] * n_planes, kinds=[0] * n_planes, n_block=[0] * n_planes, filt=None)

    def tile_put_filter(self, slot, filt):
        assert filt.size == self.slots[slot]["n"]
        self.slots[slot]["filt"] = filt

    def tile_put_bcl(self, slot, plane, data):
        s = self.slots[slot]
        assert data.size == s["n"]
        s["planes"][plane], s["kinds"][plane], s["n_block"][plane] = data, _lib.PLANE_BCL, data.size

    def tile_put_cbcl(self, slot, plane, nibbles, n_block, excluded):
        s = self.slots[slot]
        assert nibbles.size * 2 >= n_block
        s["planes"][plane], s["n_block"][plane] = nibbles, n_block
        s["kinds"][plane] = _lib.PLANE_CBCL_EXCL if excluded else _lib.PLANE_CBCL

    def sync(self):
        pass

    def _tile(self, slot, order):
        s = self.slots[slot]
        planes, kinds = [], []
        for p in order:
            k = s["kinds"][p]
            used = s["n_block"][p] if k == _lib.PLANE_BCL else (s["n_block"][p] + 1) // 2
            planes.append(np.array(s["planes"][p][:used]))
            kinds.append(KIND_NAME[k])
        return planes, kinds, np.array(s["filt"][:s["n"]])

    # -- kernels --------------------------------------------------------------------------------------
    def count(self, first_slot, n_tiles, plane_order, edit_distance=2, hamming=False, mode=0, per_target=True):
        pts, cnts, self.pairs, self.pair_codes = [], [], [], []
        self.modes.append(mode)
        self._len = len(plane_order)
        for k in range(n_tiles):
            planes, kinds, filt = self._tile(first_slot + k, plane_order)
            pt, cnt = CP.count_tile(planes, kinds, filt, self.centres, self.offs, self.idx, self.levels, edit_distance, hamming)
            pts.append(pt)
            cnts.append(cnt)
            if mode in (_lib.MODE_TWO_PASS, _lib.MODE_FUSED_LOG):
                wells = np.unique(np.concatenate([self.centres, self.idx]))
                codes, _ = CP.get_codes(planes, kinds, filt, wells.astype(np.int64))
                row = {int(w): codes[i] for i, w in enumerate(wells)}
                for t, c in enumerate(self.centres):
                    if not pt[t, 0]:
                        continue
                    for lev in range(self.levels):
                        lo, hi = self.offs[t * self.levels + lev], self.offs[t * self.levels + lev + 1]
                        for w in self.idx[lo:hi]:
                            a, b = row[int(c)], row[int(w)]
                            d = int((a != b).sum()) if hamming else CP.levenshtein(a, b)
                            if d <= edit_distance:
                                self.pairs.append((k, t, int(w), d))
                                self.pair_codes.append(np.stack([a, b]))
        return np.array(pts), np.array(cnts)

    def count_async(self, first_slot, n_tiles, plane_order, edit_distance=2, hamming=False, mode=0, want_per_target=False):
        self._pending = self.count(first_slot, n_tiles, plane_order, edit_distance, hamming, mode)

    def count_fetch(self):
        return self._pending

    def dup_pairs(self, with_seqs=False):
        rows = np.array(self.pairs, np.int32).reshape(-1, 4)
        if not with_seqs:
            return rows
        return rows, np.array(self.pair_codes, np.uint8).reshape(len(self.pairs), 2, self._len)

    def get_seqs(self, slot, indices, plane_order):
        planes, kinds, filt = self._tile(slot, plane_order)
        return CP.get_codes(planes, kinds, filt, np.asarray(indices, np.int64))


@pytest.fixture()
def oracle_engine(monkeypatch):
    eng = OracleEngine()
    st = staging.Stager(pinned=False, threads=3)
    monkeypatch.setattr(reader, "_engine", eng)
    monkeypatch.setattr(reader, "_stager", st)
    yield eng
    st.close()


@pytest.mark.parametrize("case", MAN["count"], ids=lambda c: c["name"])
def test_count_cli_host_side_matches_reference(case, oracle_engine):
    want_out, want_err = golden_count_output(case)
    out, err = run_count_case(count_cli.main, case)
    assert out == want_out
    assert err == want_err
    quiet = "-q" in case["args"] or "--quiet" in case["args"]
    # the planes stay where the inflate put them; without -q the fused kernel also logs the duplicate pairs
    assert set(oracle_engine.calls) == {"map"}
    assert set(oracle_engine.modes) == ({_lib.MODE_FUSED} if quiet else {_lib.MODE_FUSED_LOG})


def test_quiet_and_logged_runs_print_the_same_report(oracle_engine):
    """The walk with (-q) and without the duplicate-pair log gives one report."""
    case = [c for c in MAN["count"] if c["name"] == "two_lanes"][0] if any(c["name"] == "two_lanes" for c in MAN["count"]) else MAN["count"][0]
    base = ["-f", os.path.join(GOLDEN, case["targets"]), "-r", os.path.join(GOLDEN, case["run"])] + [a for a in case["args"] if a not in ("-q", "--quiet")]
    outs = []
    for extra in ([], ["-q"]):
        out, err = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            count_cli.main(base + extra)
        outs.append(out.getvalue())
        assert (err.getvalue() == "") == bool(extra)
    assert outs[0] == outs[1] and outs[0]


def test_batched_lane_with_mixed_tile_sizes_equals_the_reference_port(tmp_path, oracle_engine, monkeypatch):
    """A lane of eight tiles of two sizes with a pinned budget that holds three tiles: the -q walk
    of the smaller size (batches of 3, 1, 2, 2 host-mapped tiles) prints what the reference's loops print."""
    from oracle import ref_port as R
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(33)
    run = str(tmp_path / "run")
    row_len, ncyc = 60, 16
    sizes = [5000, 5000, 5000, 5000, 6100, 6100, 6100, 6100]
    tiles = ["11%02d" % (k + 1) for k in range(8)]
    for name, n in zip(tiles, sizes):
        td = synth.make_tile(rng, n, ncyc, row_len, pf_rate=0.7, dup_rate=0.3, shift_share=0.3, nocall_rate=0.01)
        synth.write_bcl_tile(run, 3, int(name), td, compresslevel=int(rng.integers(1, 9)))
    X, Y = synth.hex_lattice(5000, row_len)
    centres = rng.choice(np.arange(1000, 4000), size=60, replace=False)
    rings = [R.ring_indexes(X, Y, int(c)) for c in centres]
    target_file = str(tmp_path / "targets.list")
    with open(target_file, "w") as fh:
        fh.write(R.target_file_text([int(c) for c in centres], rings))
    want = R.count_run(run, target_file, "3", tiles, 5, [(2, 9), (11, 15)], edit_distance=2, verbose=True)
    per_tile = 2 * 1 + 11 * staging._round_up(5000 + 4, 256) + staging._round_up(5000, 256)
    monkeypatch.setattr(staging, "PINNED_BUDGET_BYTES", 3 * per_tile + 100)
    batches = []
    real_count = oracle_engine.count

    def spy(first_slot, n_tiles, *a, **k):
        batches.append(n_tiles)
        return real_count(first_slot, n_tiles, *a, **k)
    monkeypatch.setattr(oracle_engine, "count", spy)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        count_cli.main(["-f", target_file, "-r", run, "-s", "1108", "-i", "3", "-l", "5", "--cycles", "2-9,11-15", "-q"])
    assert out.getvalue() == want
    assert batches == [3, 1, 2, 2]          # the budget holds three 5000-well tiles or two 6100-well ones


@pytest.mark.parametrize("case", MAN["count"], ids=lambda c: c["name"])
def test_quiet_run_prints_the_reference_report_from_counter_rows(case, oracle_engine):
    """With -q the report is written straight from the counter rows of the device reduction
    (report.write_report) instead of the per-target lists: same stdout as the reference for every
    golden case -- the stdout of count_well_duplicates.py does not depend on -q -- including the lane
    without hits that ends in ZeroDivisionError after its per-tile lines."""
    want_out, _ = golden_count_output(case)
    out, err = run_count_case(count_cli.main, case, extra=[] if "-q" in case["args"] else ["-q"])
    assert out == want_out and err == ""
    assert set(oracle_engine.calls) == {"map"}


def test_quiet_lane_without_any_valid_target(tmp_path, oracle_engine):
    """No centre passes the filter anywhere in the lane: the reference's output_writer never learns the
    level count and prints the summary without level lines (count_well_duplicates.py:41-47, :124-125)."""
    from oracle import ref_port as R
    from well_duplicates_b200 import synth
    rng = np.random.default_rng(4)
    run = str(tmp_path / "run")
    for name in ("1101", "1102"):
        td = synth.make_tile(rng, 3000, 8, 50, pf_rate=0.5)
        td.filt[:] = 0
        synth.write_bcl_tile(run, 1, int(name), td)
    X, Y = synth.hex_lattice(3000, 50)
    centres = [700, 1500, 2100]
    target_file = str(tmp_path / "targets.list")
    with open(target_file, "w") as fh:
        fh.write(R.target_file_text(centres, [R.ring_indexes(X, Y, c) for c in centres]))
    want = R.count_run(run, target_file, "1", ["1101", "1102"], 3, [(0, 8)], verbose=True)
    assert "Level" not in want and "0.00%" in want
    for flags in ([], ["-q"]):
        out, err = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
            count_cli.main(["-f", target_file, "-r", run, "-s", "1102", "-i", "1", "--cycles", "0-8"] + flags)
        assert out.getvalue() == want


def test_report_text_feeds_the_wiki_formatters(tmp_path, oracle_engine):
    """The stdout of the drop-in CLI is a compatibility surface: the reference's summary_to_wiki.py /
    summary_to_wiki2.py parse it (through `tail`).  Lane reports written by count_cli -> summarize_all_lanes ->
    the formatters give exactly what the unmodified scripts made of the reference's own reports."""
    import wiki_formatters as W
    from well_duplicates_b200 import workflow as wf
    wiki = os.path.join(GOLDEN, "wiki")
    lane_files = []
    for lane, args in MAN["wiki"]["lanes"]:
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            count_cli.main(["-f", os.path.join(GOLDEN, MAN["wiki"]["targets"]), "-r", os.path.join(GOLDEN, MAN["wiki"]["run"]), "-q"] + args)
        path = tmp_path / ("40targets_lane%s.txt" % lane)
        path.write_text(out.getvalue())
        with open(os.path.join(wiki, path.name)) as fh:
            assert out.getvalue() == fh.read()
        lane_files.append(str(path))
    for tag, extra in (("all_lanes", 0), ("all_lanes_plus4", 3)):
        text = wf.summarize_all_lanes(lane_files, levels=5, extra=extra, names=[os.path.basename(p) for p in lane_files])
        for fn, ext in ((W.to_wiki, "wiki"), (W.to_wiki2, "wiki2.html")):
            with open(os.path.join(wiki, "40targets_%s.%s" % (tag, ext))) as fh:
                assert fn(text) == fh.read()


def test_empty_target_list_fails_like_the_reference(tmp_path, oracle_engine):
    """An empty target file: the reference logs the first tile and dies with IndexError in get_seqs
    (bcl_direct_reader.py:186, probed on the unmodified script) -- not with the library's ValueError."""
    empty = tmp_path / "empty.list"
    empty.write_text("")
    err = io.StringIO()
    with contextlib.redirect_stderr(err), pytest.raises(IndexError):
        count_cli.main(["-f", str(empty), "-r", os.path.join(GOLDEN, "run_bcl"), "-s", "hiseq_x", "-i", "1", "-t", "1101",
                        "-l", "5", "--cycles", "0-14"])
    assert err.getvalue() == "Reading tile 1101 in lane 1\n"


def test_truncated_plane_in_mid_lane_fails_where_the_reference_would(tmp_path, oracle_engine):
    """A .bcl.gz of the second tile is cut short.  The reference counts and logs tile 1101, logs "Reading tile 1102"
    and dies in gzip.open().read() with EOFError (bcl_direct_reader.py:207-208).  Here the three tiles are one
    batch whose inflate fails as a whole: lane_batches loads it again tile by tile, so tile 1101 is still counted
    and logged before the error surfaces at tile 1102."""
    import shutil
    run = tmp_path / "run"
    shutil.copytree(os.path.join(GOLDEN, "run_bcl"), run)
    victim = run / "Data" / "Intensities" / "BaseCalls" / "L001" / "C3.1" / "s_1_1102.bcl.gz"
    data = victim.read_bytes()
    victim.write_bytes(data[:len(data) // 2])
    case = [c for c in MAN["count"] if c["name"] == "lev_default"][0]
    _, ref_err = golden_count_output(case)
    mark = "Reading tile 1102 in lane 1\n"
    want_err = ref_err[:ref_err.index(mark) + len(mark)]
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err), pytest.raises(EOFError):
        count_cli.main(["-f", os.path.join(GOLDEN, case["targets"]), "-r", str(run), "-s", "hiseq_x", "-i", "1", "-t", "1101,1102,1103",
                        "-l", "5", "--cycles", "0-14"])
    assert out.getvalue() == "" and err.getvalue() == want_err
