"""The callers either side of the hot path: what the
reference's Snakefile.count_dups and get_cached_targets.sh do around the two
scripts, as plain Python so that a whole run is one command instead of a
Snakemake DAG.

* ``run_parameters``  -- RunInfo.xml -> last lane / last tile / lanes / cycle
  window (Snakefile.count_dups:96-100, 113-134);
* ``get_cached_targets`` -- the md5-keyed pool of target lists with its
  ``noclobber`` + ``.done`` protocol (get_cached_targets.sh:27-47,
  Snakefile.count_dups:162-186);
* ``summarize_all_lanes`` -- ``tail -n (levels+1)`` of the per-lane reports
  (Snakefile.count_dups:146-151; ``extra=3`` gives the ``+4`` variant of
  Snakefile.count_and_push:172 that keeps the duplication lines);
* ``main`` -- ``python -m well_duplicates_b200.workflow RUN [WORKDIR]``: target
  list (cached), one report per lane, the summary file -- same file names as
  the reference's rules.  Nothing here touches the GPU except through
  ``prepare_cli`` and ``count_cli``.
"""
import argparse
import contextlib
import hashlib
import io
import os
import sys
import xml.etree.ElementTree as ET

TARGETS_TO_SAMPLE = 2500      # Snakefile.count_dups:96
READ_LENGTH = 50              # :98
LEVELS_TO_SCAN = 5            # :99


def run_parameters(run_dir, read_length=READ_LENGTH):
    """RunInfo.xml -> dict(last_lane, last_tile, lanes, start_pos, end_pos).

    Tiles in the XML are ``<lane>_<tile>`` strings in no particular order; the
    reference takes the maximum string and splits it (Snakefile.count_and_push:138;
    Snakefile.count_dups:120 lacks the split and cannot run).  Read 1 longer than
    read_length + 20 cycles: the window starts at cycle 20, else at 0 (:127-133)."""
    root = ET.parse(os.path.join(run_dir, "RunInfo.xml")).getroot()
    last = max(te.text for te in root.findall(".//Tiles/Tile"))
    last_lane, last_tile = last.split("_")
    num_cycles = int(root.find("Run/Reads/Read[@Number='1']").get("NumCycles"))
    if num_cycles > read_length + 20:
        start_pos = 20
    else:
        assert num_cycles > read_length
        start_pos = 0
    return {"last_lane": last_lane, "last_tile": last_tile, "lanes": list(range(1, int(last_lane) + 1)),
            "start_pos": start_pos, "end_pos": read_length + start_pos}


def md5sum(path):
    h = hashlib.md5()
    with open(path, "rb") as fh:
        for block in iter(lambda: fh.read(1 << 22), b""):
            h.update(block)
    return h.hexdigest()


def _prepare(locs_file, target_count, out_fh):
    """prepare_cluster_indexes.py -n COUNT -f LOCS > out (stage 1 on the GPU)."""
    from . import prepare_cli
    with contextlib.redirect_stdout(out_fh):
        prepare_cli.main(["-n", str(target_count), "-f", locs_file])


def get_cached_targets(locs_file, target_count, output_file, cluster_lists=None, prepare=_prepare):
    """get_cached_targets.sh: the output is a symlink into the cache when the
    cache directory exists, a regular file otherwise; it is never clobbered.
    Two processes that try to fill the same cache entry: the second fails on
    the exclusive create (``set -o noclobber``); a failed prepare removes its
    partial entry (the script's EXIT trap); ``.done`` marks a complete one."""
    if os.path.lexists(output_file):
        raise FileExistsError(output_file)
    if cluster_lists is None:
        cluster_lists = os.environ.get("CLUSTER_LISTS", os.path.join(os.path.dirname(os.path.abspath(__file__)), "cluster_lists"))
    if os.path.isdir(cluster_lists):
        cached = os.path.join(cluster_lists, "%sclusters_%s.list" % (target_count, md5sum(locs_file)))
        if not os.path.exists(cached + ".done"):
            with open(cached, "x") as fh:
                try:
                    prepare(locs_file, target_count, fh)
                except BaseException:
                    fh.close()
                    os.remove(cached)
                    raise
            open(cached + ".done", "a").close()
        if os.path.isabs(cached):
            os.symlink(cached, output_file)
        else:
            os.symlink(os.path.relpath(cached, os.path.dirname(os.path.abspath(output_file))), output_file)
        return cached
    with open(output_file, "x") as fh:
        prepare(locs_file, target_count, fh)
    return output_file


def summarize_all_lanes(lane_files, levels=LEVELS_TO_SCAN, extra=0, names=None):
    """What ``tail -n (levels + 1 + extra) file...`` prints (GNU tail: a
    ``==> name <==`` header per file when there are several, a blank line
    between files).  ``names``: the file names as the command line would have
    spelled them (the reference runs inside its working directory)."""
    n = levels + 1 + extra
    names = lane_files if names is None else names
    out = io.StringIO()
    for k, path in enumerate(lane_files):
        with open(path) as fh:
            lines = fh.readlines()
        if len(lane_files) > 1:
            out.write("%s==> %s <==\n" % ("\n" if k else "", names[k]))
        out.write("".join(lines[-n:] if n else []))
    return out.getvalue()


def main(argv=None):
    ap = argparse.ArgumentParser(description="well_duplicates for every lane of a run (what Snakefile.count_dups does)")
    ap.add_argument("run", help="sequencer output directory (RunInfo.xml, Data/Intensities/...)")
    ap.add_argument("workdir", nargs="?", default=".")
    ap.add_argument("-n", "--targets", type=int, default=TARGETS_TO_SAMPLE)
    ap.add_argument("-l", "--levels", type=int, default=LEVELS_TO_SCAN)
    ap.add_argument("--read-length", type=int, default=READ_LENGTH)
    ap.add_argument("-S", "--summary-only", action="store_true")
    ap.add_argument("--cluster-lists", default=None, help="target-list cache directory (default $CLUSTER_LISTS)")
    args = ap.parse_args(argv)
    from . import count_cli
    rp = run_parameters(args.run, args.read_length)
    os.makedirs(args.workdir, exist_ok=True)
    targfile = os.path.join(args.workdir, "%dclusters.list" % args.targets)
    if not os.path.lexists(targfile):
        get_cached_targets(os.path.join(args.run, "Data", "Intensities", "s.locs"), args.targets, targfile,
                           cluster_lists=args.cluster_lists)
    lane_files = []
    for lane in rp["lanes"]:
        path = os.path.join(args.workdir, "%dtargets_lane%d.txt" % (args.targets, lane))
        cmd = ["-f", targfile, "-n", str(args.targets), "-s", rp["last_tile"], "-r", args.run, "-i", str(lane),
               "-l", str(args.levels), "--cycles", "%d-%d" % (rp["start_pos"], rp["end_pos"])]
        if args.summary_only:
            cmd.append("-S")
        with open(path, "w") as fh, contextlib.redirect_stdout(fh):
            count_cli.main(cmd)
        lane_files.append(path)
    summary = os.path.join(args.workdir, "%dtargets_all_lanes.txt" % args.targets)
    with open(summary, "w") as fh:
        fh.write(summarize_all_lanes(lane_files, args.levels, names=[os.path.basename(p) for p in lane_files]))
    sys.stderr.write("wrote %s\n" % summary)


if __name__ == "__main__":
    main()
