"""Shared helpers for the test-suite (no GPU needed to import)."""
import json
import os

import fixture_inputs as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


def locs_path(name, tmpdir=None):
    """Path of a locs fixture; the large generated ones are rebuilt on demand."""
    from well_duplicates_b200 import synth
    p = os.path.join(GOLDEN, "locs", name + ".locs")
    if os.path.exists(p):
        return p
    assert tmpdir is not None, "fixture %s is generated; pass a tmp dir" % name
    p = os.path.join(str(tmpdir), name + ".locs")
    if not os.path.exists(p):
        synth.write_locs(p, fx.LOCS_FIXTURES[name]())
    return p


def parse_count_args(args):
    """argv of a golden count case -> dict of the options that matter."""
    o = {"levels": 3, "edit": 2, "hamming": False, "summary": False, "limit": 2500,
         "lanes": None, "tiles": None, "stype": None, "x": 50, "y": 100, "cycles": None, "quiet": False}
    it = iter(args)
    for a in it:
        if a == "-l":
            o["levels"] = int(next(it))
        elif a == "-e":
            o["edit"] = int(next(it))
        elif a == "--hamming":
            o["hamming"] = True
        elif a == "-S":
            o["summary"] = True
        elif a == "-q":
            o["quiet"] = True
        elif a == "-n":
            o["limit"] = int(next(it))
        elif a == "-i":
            o["lanes"] = next(it)
        elif a == "-t":
            o["tiles"] = next(it)
        elif a == "-s":
            o["stype"] = next(it)
        elif a == "-x":
            o["x"] = int(next(it))
        elif a == "-y":
            o["y"] = int(next(it))
        elif a == "--cycles":
            o["cycles"] = next(it)
        else:
            raise ValueError(a)
    if o["cycles"]:
        o["ranges"] = [tuple(int(v) for v in r.split("-")) for r in o["cycles"].split(",")]
    else:
        o["ranges"] = [(o["x"], o["y"])]
    return o


_EXC = {"ZeroDivisionError": ZeroDivisionError, "FileNotFoundError": FileNotFoundError, "RuntimeError": RuntimeError,
        "IndexError": IndexError, "AssertionError": AssertionError}


def run_count_case(main, case, extra=()):
    """Runs a count CLI ``main(argv)`` on a golden case -> (stdout, stderr); a case the reference dies on must
    die here with the same exception type (manifest "raises")."""
    import contextlib
    import io

    import pytest
    argv = ["-f", os.path.join(GOLDEN, case["targets"]), "-r", os.path.join(GOLDEN, case["run"])] + case["args"] + list(extra)
    out, err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        if case["returncode"] != 0:
            with pytest.raises(_EXC[case.get("raises") or "ZeroDivisionError"]):
                main(argv)
        else:
            main(argv)
    return out.getvalue(), err.getvalue()


def golden_count_output(case):
    """(stdout, log) of the unmodified reference; for a run that died the log ends where its traceback began."""
    with open(os.path.join(GOLDEN, "count", case["name"] + ".stdout")) as fh:
        out = fh.read()
    with open(os.path.join(GOLDEN, "count", case["name"] + ".stderr")) as fh:
        err = fh.read()
    if case["returncode"] != 0:
        err = err[:err.rstrip("\n").rfind("\n") + 1]          # drop the exception line
    return out, err
