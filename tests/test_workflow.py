"""Host logic around the hot path: RunInfo.xml parameters, the
md5-keyed target-list cache, the all-lanes summary.  No GPU: the prepare step
is injected."""
import os
import subprocess

import pytest

from well_duplicates_b200 import workflow as wf

RUNINFO = """<?xml version="1.0"?>
<RunInfo Version="4">
  <Run Id="r" Number="1">
    <Reads>
      <Read Number="1" NumCycles="%d" IsIndexedRead="N" />
      <Read Number="2" NumCycles="8" IsIndexedRead="Y" />
    </Reads>
    <FlowcellLayout LaneCount="2" SurfaceCount="2" SwathCount="2" TileCount="28">
      <TileSet><Tiles>
        <Tile>1_2128</Tile><Tile>2_2228</Tile><Tile>1_1101</Tile><Tile>2_1101</Tile><Tile>1_2228</Tile>
      </Tiles></TileSet>
    </FlowcellLayout>
  </Run>
</RunInfo>
"""


def _run(tmp_path, cycles):
    d = tmp_path / ("run%d" % cycles)
    d.mkdir()
    (d / "RunInfo.xml").write_text(RUNINFO % cycles)
    return str(d)


def test_run_parameters_window_and_last_tile(tmp_path):
    # Snakefile.count_dups:113-134 (with the split of Snakefile.count_and_push:138)
    rp = wf.run_parameters(_run(tmp_path, 151))
    assert rp == {"last_lane": "2", "last_tile": "2228", "lanes": [1, 2], "start_pos": 20, "end_pos": 70}
    rp = wf.run_parameters(_run(tmp_path, 51))
    assert (rp["start_pos"], rp["end_pos"]) == (0, 50)
    assert wf.run_parameters(_run(tmp_path, 71))["start_pos"] == 20
    assert wf.run_parameters(_run(tmp_path, 70))["start_pos"] == 0
    with pytest.raises(AssertionError):
        wf.run_parameters(_run(tmp_path, 50))


def test_target_cache_protocol(tmp_path):
    """get_cached_targets.sh:27-47: md5-keyed entry, noclobber, .done sentinel, symlink out."""
    locs = tmp_path / "s.locs"
    locs.write_bytes(b"\x01\x00\x00\x00\x00\x00\x80\x3f\x02\x00\x00\x00" + b"\0" * 16)
    cache = tmp_path / "cluster_lists"
    cache.mkdir()
    calls = []

    def prepare(path, n, fh):
        calls.append((path, n))
        fh.write("7\n1,2\n")

    out1 = tmp_path / "a.list"
    cached = wf.get_cached_targets(str(locs), 2500, str(out1), str(cache), prepare)
    md5 = subprocess.run(["md5sum", str(locs)], capture_output=True, text=True).stdout.split()[0]
    assert cached == str(cache / ("2500clusters_%s.list" % md5))
    assert os.path.islink(out1) and out1.read_text() == "7\n1,2\n" and os.path.exists(cached + ".done")
    # second request: served from the cache, prepare is not run again
    out2 = tmp_path / "b.list"
    wf.get_cached_targets(str(locs), 2500, str(out2), str(cache), prepare)
    assert len(calls) == 1 and out2.read_text() == "7\n1,2\n"
    # a different target count is a different entry
    wf.get_cached_targets(str(locs), 10, str(tmp_path / "c.list"), str(cache), prepare)
    assert len(calls) == 2
    # output is never clobbered
    with pytest.raises(FileExistsError):
        wf.get_cached_targets(str(locs), 2500, str(out1), str(cache), prepare)
    # an entry somebody else is still writing (no .done): fail, do not overwrite
    half = cache / ("77clusters_%s.list" % md5)
    half.write_text("partial")
    with pytest.raises(FileExistsError):
        wf.get_cached_targets(str(locs), 77, str(tmp_path / "d.list"), str(cache), prepare)
    assert half.read_text() == "partial"

    # a failing prepare leaves neither a partial entry nor a sentinel
    def broken(path, n, fh):
        fh.write("1\n")
        raise RuntimeError("Got no wells for cluster 1 level 0")
    with pytest.raises(RuntimeError):
        wf.get_cached_targets(str(locs), 5, str(tmp_path / "e.list"), str(cache), broken)
    assert not os.path.exists(cache / ("5clusters_%s.list" % md5)) and not os.path.lexists(tmp_path / "e.list")
    # no cache directory: a regular file
    out3 = tmp_path / "f.list"
    wf.get_cached_targets(str(locs), 2500, str(out3), str(tmp_path / "nowhere"), prepare)
    assert not os.path.islink(out3) and out3.read_text() == "7\n1,2\n"


def test_relative_cache_gives_relative_link(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    os.mkdir("cl")
    os.mkdir("work")
    open("s.locs", "wb").write(b"x")
    wf.get_cached_targets("s.locs", 3, "work/3clusters.list", "cl", lambda p, n, fh: fh.write("1\n2,3\n"))
    assert not os.path.isabs(os.readlink("work/3clusters.list"))
    assert open("work/3clusters.list").read() == "1\n2,3\n"


@pytest.mark.parametrize("n_files,levels,extra", [(1, 5, 0), (3, 5, 0), (2, 3, 3), (2, 5, 40)])
def test_summary_is_what_tail_prints(tmp_path, monkeypatch, n_files, levels, extra):
    """Snakefile.count_dups:146-151 / Snakefile.count_and_push:172."""
    monkeypatch.chdir(tmp_path)
    names = []
    for k in range(n_files):
        name = "2500targets_lane%d.txt" % (k + 1)
        with open(name, "w") as fh:
            fh.write("".join("Lane %d line %d\n" % (k + 1, i) for i in range(12 + k)))
        names.append(name)
    want = subprocess.run(["tail", "-n", str(levels + 1 + extra)] + names, capture_output=True, text=True).stdout
    assert wf.summarize_all_lanes(names, levels, extra) == want


# ---- the consumers of the report: GNU tail summary and the wiki formatters -------------------------
WIKI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wiki")


def _read(name):
    with open(os.path.join(WIKI, name)) as fh:
        return fh.read()


@pytest.mark.parametrize("tag,extra", [("all_lanes", 0), ("all_lanes_plus4", 3)])
def test_summary_equals_gnu_tail_of_the_reference_lane_files(tag, extra):
    """summarize_all_lanes on the reference's own per-lane reports == what `tail -n` printed for them
    (Snakefile.count_dups:151: levels + 1 lines; Snakefile.count_and_push:172: levels + 4)."""
    lanes = ["40targets_lane1.txt", "40targets_lane2.txt"]
    got = wf.summarize_all_lanes([os.path.join(WIKI, f) for f in lanes], levels=5, extra=extra, names=lanes)
    assert got == _read("40targets_%s.txt" % tag)


@pytest.mark.parametrize("tag", ["all_lanes", "all_lanes_plus4"])
def test_wiki_formatter_restatements_are_pinned_to_the_reference_scripts(tag):
    import wiki_formatters as W
    text = _read("40targets_%s.txt" % tag)
    assert W.to_wiki(text) == _read("40targets_%s.wiki" % tag)
    assert W.to_wiki2(text) == _read("40targets_%s.wiki2.html" % tag)
