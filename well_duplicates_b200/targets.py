"""Target lists: the text format written by prepare_cluster_indexes and the
containers count_well_duplicates iterates over.

Mirrors the interface of the reference's target.py (load_targets :6-40,
AllTargets :42-106, Target :108-144) -- same names, arguments, iteration order
and exceptions -- and adds ``AllTargets.to_csr`` which is what the CUDA library
consumes (wd_targets_load)."""
import numpy as np


class Target:
    """One sampled well plus its rings: ``coords[0] == [centre]``,
    ``coords[k]`` = wells of ring k."""

    def __init__(self, coords):
        assert len(coords[0]) == 1, "Centre of target must be a single int, not " + str(coords)
        self.coords = coords

    def get_indices(self, level=None):
        if level is None:
            return [w for ring in self.coords for w in ring]
        return self.coords[level]

    def get_centre(self):
        return self.coords[0][0]

    def get_levels(self):
        return len(self.coords)

    def get_level_from_index(self, index):
        for lev, ring in enumerate(self.coords):
            if index in ring:
                return lev
        return None


class AllTargets:
    def __init__(self):
        self._by_centre = {}
        self._by_well = {}
        self.levels = None

    def __len__(self):
        return len(self._by_centre)

    def __iter__(self):
        return iter(self._by_centre.values())

    def get_target_by_centre(self, centre):
        return self._by_centre[centre]

    def add_target(self, coords):
        tgt = Target(coords)
        centre = tgt.get_centre()
        assert centre not in self._by_centre          # target.py:72
        if self.levels is None:
            self.levels = tgt.get_levels()
        else:
            assert self.levels == tgt.get_levels()    # target.py:75-78
        self._by_centre[centre] = tgt
        for w in tgt.get_indices():
            self._by_well.setdefault(w, []).append(tgt)

    def get_all_indices(self, level=None):
        if level == 0:
            return list(self._by_centre)
        if level is None:
            return list(self._by_well)
        return [w for tgt in self for w in tgt.get_indices(level)]

    def get_from_index(self, index):
        return [(tgt, tgt.get_level_from_index(index)) for tgt in self._by_well.get(index, [])]

    def to_csr(self, rings=None):
        """(centres[t], level_offsets[t*rings+1], idx) as uint32 arrays, rings
        1..``rings`` of every target in iteration (file) order."""
        have = (self.levels or 1) - 1
        rings = have if rings is None else rings
        if rings > have:
            raise IndexError("list index out of range")     # what Target.get_indices(level) would raise
        centres = np.fromiter((t.get_centre() for t in self), dtype=np.int64, count=len(self))
        lens = np.fromiter((len(t.coords[k]) for t in self for k in range(1, rings + 1)), dtype=np.int64,
                           count=len(self) * rings)
        offs = np.zeros(lens.size + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        idx = np.fromiter((w for t in self for k in range(1, rings + 1) for w in t.coords[k]), dtype=np.int64,
                          count=int(offs[-1]))
        for arr in (centres, idx):
            if arr.size and (arr.min() < 0 or arr.max() >= 2 ** 32):
                bad = int(arr.min() if arr.min() < 0 else arr.max())
                if bad < 0:
                    raise IndexError("Requested cluster %i is a negative number." % bad)
                raise IndexError("Requested cluster %i is out of range." % bad)
        return centres.astype(np.uint32), offs.astype(np.uint32), idx.astype(np.uint32)


def load_targets(filename, levels=None, limit=None):
    """Reads a target file.  A line without a comma starts a new record; the
    first ``levels`` lines of each record are kept (centre included); reading
    stops after ``limit`` records."""
    out = AllTargets()

    def flush(rec):
        out.add_target([[int(v) for v in line.split(",")] for line in rec[:levels]])

    with open(filename, "r") as fh:
        rec = None
        for raw in fh:
            line = raw.rstrip()
            if "," in line:
                rec.append(line)        # AttributeError on a file that starts with a ring line, as in the reference
                continue
            if rec:
                flush(rec)
                if limit and len(out) == limit:
                    return out
            rec = [line]
        if rec:
            flush(rec)
    return out
