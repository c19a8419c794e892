"""Command line of prepare_cluster_indexes.py (reference :81-97, :121-169) on
the GPU: samples wells with Python's own ``random.sample`` (so the sample is
identical by construction), then builds every target's rings with the
spatial-hash kernels (K0-K2) and prints the target file."""
import random
import struct
import sys
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np

from .engine import Engine
from .reader import default_engine

__VERSION__ = 0.2
DEF_SAMPLE_SIZE = 2500
LEVELS = 5                        # len(MAX_DISTS) - 1


def log(msg):
    print(str(msg), file=sys.stderr)


def parse_args(argv=None):
    p = ArgumentParser(description="Picks n random wells from an s.locs file and lists the wells in the 1st to "
                                   "5th ring around each.",
                       formatter_class=ArgumentDefaultsHelpFormatter)
    p.add_argument("-f", "--slocs", dest="slocs", type=str, required=True, help="the s.locs file")
    p.add_argument("-s", "--seed", dest="seed", type=int, default=None, help="seed for the random well selection")
    p.add_argument("-n", "--sample_size", dest="sample_size", type=int, default=DEF_SAMPLE_SIZE,
                   help="number of wells to sample")
    # extension (not in the reference): the same list as arrays (targets.py), for samples too large for text
    p.add_argument("--binary", metavar="FILE", help="write the target list to FILE in the binary format that "
                   "count_well_duplicates.py -f also accepts, instead of printing it")
    return p.parse_args(argv)


def read_locs(path):
    """-> (cluster count from the header, float32 [n, 2])."""
    with open(path, "rb") as fh:
        n = int(struct.unpack("=ifI", fh.read(12))[2])
        body = fh.read()
    xy = np.frombuffer(body, dtype="<f4", count=(len(body) // 8) * 2).reshape(-1, 2)
    return n, xy


def get_random_array(r_max, r_l, seed):
    if seed:                       # 0 / None leave the generator unseeded, as in the reference (:26-30)
        random.seed(seed)
    return random.sample(range(r_max), r_l)


def format_targets(centres, level_offsets, idx, levels=LEVELS):
    lines = []
    for t, c in enumerate(centres):
        lines.append(str(int(c)))
        for l in range(levels):
            a, b = level_offsets[t * levels + l], level_offsets[t * levels + l + 1]
            lines.append(",".join(map(str, idx[a:b].tolist())))
    return "".join(x + "\n" for x in lines)


def build_targets(xy, centres, engine=None, as_arrays=False):
    """Rings for ``centres`` -> text of the target file (or the CSR arrays).  Raises RuntimeError
    (reference :70-76) when a ring is empty."""
    eng = engine if engine is not None else default_engine()
    eng.load_locs(xy)
    try:
        offs, idx = eng.ring_query(centres, LEVELS)
    except RuntimeError as err:
        fe = getattr(err, "first_empty", None)
        if fe is None:
            raise
        x, y = eng.pixels()
        c = int(centres[fe[0]])
        failure = RuntimeError("Got no wells for cluster %s at (%s,%s) level %s", (c, int(x[c]), int(y[c]), fe[1]))
        failure.target_ordinal = fe[0]
        raise failure
    if as_arrays:
        return offs, idx
    return format_targets(centres, offs, idx)


def main(argv=None):
    args = parse_args(argv)
    log("seed: %s" % (args.seed))
    log("sample size: %s" % (args.sample_size))
    n, xy = read_locs(args.slocs)
    log("Maximum number of cluster according to s.locs: %s" % n)
    sample = get_random_array(n, args.sample_size, args.seed)
    log(sample)
    if len(sample) and max(sample) >= xy.shape[0]:
        raise struct.error("unpack requires a buffer of 8 bytes")     # header promises more records than the file holds
    # get_indexes() logs the byte offset it seeks to for every centre it gets to (prepare_cluster_indexes.py:52-53)
    def log_offsets(upto):
        for c in sample[:upto]:
            log(max(0, c - 20000) * 8 + 12)
    try:
        offs, idx = build_targets(xy, sample, as_arrays=True)
    except RuntimeError as err:
        log_offsets(getattr(err, "target_ordinal", -1) + 1)
        raise
    log_offsets(len(sample))
    if args.binary:
        from .targets import save_targets_binary
        save_targets_binary(args.binary, sample, offs, idx, LEVELS)
        return
    sys.stdout.write(format_targets(sample, offs, idx))


if __name__ == "__main__":
    main()
