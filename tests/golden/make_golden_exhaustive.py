#!/usr/bin/env python3
"""Golden outputs for exhaustive mode (every well a target), from the UNMODIFIED
reference.  Authoring container only (needs /root/reference):

    python tests/golden/make_golden_exhaustive.py

prepare_cluster_indexes.py is asked for as many targets as the tile has wells
(-n 3072 on the committed 3072-well lattices), and count_well_duplicates.py
counts the committed synthetic run tests/golden/run_bcl against that list.
Only the reports are stored (the 1 MB target lists are not); the tests rebuild
the neighbourhoods themselves and must arrive at the same report.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
ENV = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "levenshtein_shim"), PYTHONWARNINGS="ignore")

CASES = [
    # name, locs fixture, extra argv of count_well_duplicates.py
    ("hex_small_lev", "hex_small", ["-l", "5", "--cycles", "0-14"]),
    ("hex_small_hamming", "hex_small", ["-l", "5", "--cycles", "0-14", "--hamming"]),
    ("hex_small_l3_e3", "hex_small", ["-l", "3", "-e", "3", "--cycles", "2-14"]),
    ("hex_shuffled_lev", "hex_shuffled", ["-l", "5", "--cycles", "0-14"]),
]


def main():
    out_dir = os.path.join(HERE, "exhaustive")
    os.makedirs(out_dir, exist_ok=True)
    manifest = []
    lists = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, locs, extra in CASES:
            if locs not in lists:
                p = subprocess.run([sys.executable, os.path.join(REF, "prepare_cluster_indexes.py"), "-f",
                                    os.path.join(HERE, "locs", locs + ".locs"), "-n", "3072", "-s", "1"],
                                   capture_output=True, text=True, env=ENV)
                assert p.returncode == 0, p.stderr[-2000:]
                lists[locs] = os.path.join(tmp, locs + ".list")
                with open(lists[locs], "w") as fh:
                    fh.write(p.stdout)
                print("prepare", locs, p.stdout.count("\n"), "lines")
            argv = ["-f", lists[locs], "-n", "3072", "-r", os.path.join(HERE, "run_bcl"), "-s", "hiseq_x", "-i", "1",
                    "-t", "1101,1102", "-q"] + extra
            p = subprocess.run([sys.executable, os.path.join(REF, "count_well_duplicates.py")] + argv,
                               capture_output=True, text=True, env=ENV)
            assert p.returncode == 0, p.stderr[-2000:]
            with open(os.path.join(out_dir, name + ".stdout"), "w") as fh:
                fh.write(p.stdout)
            manifest.append({"name": name, "locs": locs, "args": extra, "tiles": ["1101", "1102"], "lane": "1"})
            print("count", name, p.stdout.splitlines()[:2])
    with open(os.path.join(out_dir, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)


if __name__ == "__main__":
    main()
