#!/usr/bin/env python3
"""Regenerate tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (``/root/reference`` does not exist on the
GPU box):

    python tests/golden/make_golden.py

It writes small synthetic inputs (from well_duplicates_b200.synth and
tests/fixture_inputs.py), runs the reference scripts as subprocesses with the
``Levenshtein`` stand-in from oracle/levenshtein_shim on PYTHONPATH (the only
thing the reference needs that is not installed here), and stores what they
print.  Nothing in ``tests/`` reads /root/reference at run time.
"""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import fixture_inputs as fx  # noqa: E402
from well_duplicates_b200 import synth  # noqa: E402

ENV = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "levenshtein_shim"),
           PYTHONWARNINGS="ignore")


def run_ref(script, args, cwd=None):
    p = subprocess.run([sys.executable, os.path.join(REF, script)] + [str(a) for a in args],
                       capture_output=True, text=True, env=ENV, cwd=cwd)
    return p


def strip_warnings(stderr):
    """Python 3.12 prints SyntaxWarnings for the reference's '\\d' literals."""
    out = []
    skip = 0
    for line in stderr.splitlines(keepends=True):
        if skip:
            skip -= 1
            continue
        if "SyntaxWarning" in line:
            skip = 1
            continue
        out.append(line)
    return "".join(out)


def main():
    manifest = {"prepare": [], "count": [], "getseqs": []}

    # ---- reference's own test fixtures ---------------------------------
    rt = os.path.join(HERE, "ref_tests")
    os.makedirs(rt, exist_ok=True)
    for f in ("small.list", "bad1.list", "bad2.list"):
        shutil.copyfile(os.path.join(REF, "test", f), os.path.join(rt, f))
    code = (
        "import json, io, sys, contextlib\n"
        "sys.path.insert(0, %r)\n"
        "import test.test_count_well_duplicates as T\n"
        "from count_well_duplicates import output_writer\n"
        "cases = {\n"
        " 'full': dict(lane=1, sample_size=4, lane_dupl=T.LANE_DUPL, verbose=1, levels=0, expected=T.EXPECTED_OUT_1, sl=[0, None]),\n"
        " 'badlane_full': dict(lane=1, sample_size=4, lane_dupl=T.BAD_TILE_LANE, verbose=1, levels=0, expected=T.EXPECTED_OUT_2, sl=[0, None]),\n"
        " 'badlane_brief': dict(lane=1, sample_size=4, lane_dupl=T.BAD_TILE_LANE, verbose=0, levels=0, expected=T.EXPECTED_OUT_2, sl=[-4, None]),\n"
        " 'limited_levels': dict(lane=1, sample_size=4, lane_dupl=T.LANE_DUPL, verbose=1, levels=2, expected=T.EXPECTED_OUT_3, sl=[0, None]),\n"
        " 'empty_data': dict(lane=1, sample_size=4, lane_dupl={'1222': []}, verbose=1, levels=0, expected=T.EXPECTED_OUT_4, sl=[0, None]),\n"
        "}\n"
        "for k, c in cases.items():\n"
        "    buf = io.StringIO()\n"
        "    with contextlib.redirect_stdout(buf):\n"
        "        output_writer(c['lane'], c['sample_size'], c['lane_dupl'], verbose=c['verbose'], levels=c['levels'])\n"
        "    c['printed'] = buf.getvalue()\n"
        "json.dump(cases, sys.stdout, indent=1)\n" % REF)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=ENV, cwd=REF)
    assert p.returncode == 0, p.stderr
    with open(os.path.join(rt, "output_writer_cases.json"), "w") as fh:
        fh.write(p.stdout)

    # ---- stage 1: target files -----------------------------------------
    locs_dir = os.path.join(HERE, "locs")
    os.makedirs(locs_dir, exist_ok=True)
    for name, maker in fx.LOCS_FIXTURES.items():
        xy = maker()
        path = os.path.join(locs_dir, name + ".locs")
        synth.write_locs(path, xy)
    for name, n, seed in fx.PREPARE_CASES:
        path = os.path.join(locs_dir, name + ".locs")
        args = ["-f", path, "-n", n]
        if seed is not None:
            args += ["-s", seed]
        p = run_ref("prepare_cluster_indexes.py", args)
        tag = "%s_n%d_s%s" % (name, n, seed)
        out = os.path.join(locs_dir, tag + ".list")
        with open(out, "w") as fh:
            fh.write(p.stdout)
        # the chatter on stderr (seed, sample, the byte offset of every scan); a run that dies: up to its traceback
        perr = strip_warnings(p.stderr)
        if p.returncode != 0:
            perr = perr.split("Traceback (most recent call last)")[0]
        with open(os.path.join(locs_dir, tag + ".stderr"), "w") as fh:
            fh.write(perr)
        manifest["prepare"].append({"locs": name, "n": n, "seed": seed, "returncode": p.returncode,
                                    "list": os.path.relpath(out, HERE),
                                    "error": ("RuntimeError" if "RuntimeError" in p.stderr else None)})
        print("prepare", tag, "rc", p.returncode, "bytes", len(p.stdout))
    # big generated-on-the-fly locs are not committed
    for name in fx.LOCS_NOT_COMMITTED:
        os.remove(os.path.join(locs_dir, name + ".locs"))

    # ---- stage 2/3 inputs -----------------------------------------------
    rng = np.random.default_rng(20261018)
    row_len, n_wells, n_cyc = fx.SMALL_ROW_LEN, fx.SMALL_WELLS, fx.SMALL_CYCLES
    run_bcl = os.path.join(HERE, "run_bcl")
    run_cbcl = os.path.join(HERE, "run_cbcl")
    for d in (run_bcl, run_cbcl):
        shutil.rmtree(d, ignore_errors=True)
    tiles = {}
    for tile, dup in ((1101, 0.5), (1102, 0.35), (1103, 0.0)):
        td = synth.make_tile(rng, n_wells, n_cyc, row_len, pf_rate=0.7, nocall_rate=0.02,
                             dup_rate=dup, shift_share=0.3)
        tiles[tile] = td
        synth.write_bcl_tile(run_bcl, 1, tile, td, compresslevel=9)
    # second lane, one tile, for the lane loop
    td = synth.make_tile(rng, n_wells, n_cyc, row_len, pf_rate=0.6, nocall_rate=0.01, dup_rate=0.4)
    synth.write_bcl_tile(run_bcl, 2, 1101, td, compresslevel=9)
    ctiles = {}
    for tile, dup in ((1101, 0.5), (1102, 0.4), (2101, 0.3)):
        ctiles[tile] = synth.make_tile(rng, n_wells + (1 if tile == 1102 else 0), n_cyc, row_len,
                                       pf_rate=0.65, nocall_rate=0.02, dup_rate=dup, shift_share=0.3)
    # tile 1102 has an odd well count so the nibble padding is exercised; it
    # needs its own tile-size-compatible targets, so only get_seqs covers it.
    synth.write_cbcl_lane(run_cbcl, 1, ctiles, excluded_from_cycle=6, compresslevel=9)

    targets = os.path.join("locs", "hex_small_n40_s13.list")
    tf = os.path.join(HERE, targets)

    # ---- stage 3: count_well_duplicates ----------------------------------
    out_dir = os.path.join(HERE, "count")
    shutil.rmtree(out_dir, ignore_errors=True)
    os.makedirs(out_dir)
    for name, run, extra in fx.COUNT_CASES:
        args = ["-f", tf, "-r", os.path.join(HERE, run)] + list(extra)
        p = run_ref("count_well_duplicates.py", args)
        err = strip_warnings(p.stderr)
        with open(os.path.join(out_dir, name + ".stdout"), "w") as fh:
            fh.write(p.stdout)
        # a run that dies: the log up to the traceback (whose paths are the container's), then the exception line
        raises = None
        if p.returncode != 0:
            raises = err.strip().splitlines()[-1].split(":")[0]
            err = err.split("Traceback (most recent call last)")[0] + err.strip().splitlines()[-1].split(":")[0] + "\n"
        with open(os.path.join(out_dir, name + ".stderr"), "w") as fh:
            fh.write(err)
        manifest["count"].append({"name": name, "run": run, "args": list(map(str, extra)),
                                  "targets": targets, "returncode": p.returncode, "raises": raises})
        print("count", name, "rc", p.returncode, "stdout lines", p.stdout.count("\n"))

    # ---- the callers behind the report: all-lanes summary (GNU tail) and the two wiki formatters ----------
    wiki_dir = os.path.join(HERE, "wiki")
    shutil.rmtree(wiki_dir, ignore_errors=True)
    os.makedirs(wiki_dir)
    lane_files = []
    for lane, extra in fx.WIKI_LANES:
        p = run_ref("count_well_duplicates.py", ["-f", tf, "-r", run_bcl, "-q"] + list(extra))
        assert p.returncode == 0, p.stderr
        lane_files.append("40targets_lane%s.txt" % lane)
        with open(os.path.join(wiki_dir, lane_files[-1]), "w") as fh:
            fh.write(p.stdout)
    for tag, n in (("all_lanes", 5 + 1), ("all_lanes_plus4", 5 + 4)):       # Snakefile.count_dups:151, Snakefile.count_and_push:172
        p = subprocess.run(["tail", "-n", str(n)] + lane_files, capture_output=True, text=True, cwd=wiki_dir)
        assert p.returncode == 0, p.stderr
        with open(os.path.join(wiki_dir, "40targets_%s.txt" % tag), "w") as fh:
            fh.write(p.stdout)
        for script, ext in (("summary_to_wiki.py", "wiki"), ("summary_to_wiki2.py", "wiki2.html")):
            q = subprocess.run([sys.executable, os.path.join(REF, script)], input=p.stdout, capture_output=True, text=True, env=ENV)
            assert q.returncode == 0, q.stderr
            with open(os.path.join(wiki_dir, "40targets_%s.%s" % (tag, ext)), "w") as fh:
                fh.write(q.stdout)
    manifest["wiki"] = {"lanes": [[lane, list(map(str, extra))] for lane, extra in fx.WIKI_LANES], "run": "run_bcl",
                        "targets": targets, "levels": 5}
    print("wiki", sorted(os.listdir(wiki_dir)))

    # ---- stage 2: get_seqs ------------------------------------------------
    gs_dir = os.path.join(HERE, "getseqs")
    shutil.rmtree(gs_dir, ignore_errors=True)
    os.makedirs(gs_dir)
    for name, run, lane, tile, idx, start, end in fx.getseqs_cases():
        code = (
            "import json, sys\n"
            "import bcl_direct_reader as B\n"
            "t = B.BCLReader(%r).get_tile(%r, %r)\n"
            "try:\n"
            "    r = t.get_seqs(%r, %r, %r)\n"
            "    print(json.dumps({'ok': {str(k): [v[0], bool(v[1])] for k, v in r.items()}}))\n"
            "except Exception as e:\n"
            "    print(json.dumps({'error': type(e).__name__}))\n"
            % (os.path.join(HERE, run), lane, tile, idx, start, end))
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=ENV, cwd=REF)
        assert p.returncode == 0, p.stderr
        with open(os.path.join(gs_dir, name + ".json"), "w") as fh:
            fh.write(p.stdout)
        manifest["getseqs"].append({"name": name, "run": run, "lane": lane, "tile": tile,
                                    "indices": idx, "start": start, "end": end})
        print("getseqs", name, p.stdout[:60].strip())

    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)


if __name__ == "__main__":
    main()
