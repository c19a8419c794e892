"""Target lists: the text format written by prepare_cluster_indexes and the
containers count_well_duplicates iterates over.

Mirrors the interface of the reference's target.py (load_targets :6-40,
AllTargets :42-106, Target :108-144) -- same names, arguments, iteration order
and exceptions -- and adds ``AllTargets.to_csr`` which is what the CUDA library
consumes (wd_targets_load).

Beside the text format there is a binary one for lists the text form is too bulky for -- every
well of a tile as a target is 3 GB of text, 1.6 GB as arrays that are mapped, not parsed:

    offset  0  8 bytes   magic  b"WDTGTS\x01\x00"
            8  uint32    rings per target (the text form's level lines; 5 from prepare_cluster_indexes)
           12  uint32    0
           16  uint64    t = number of targets
           24  uint64    n = number of ring wells
           32  uint32[t]             centres, in file order
               uint64[t * rings + 1] level_offsets into idx (ring l of target i = idx[off[i*rings+l] : off[i*rings+l+1]])
               uint32[n]             idx, ascending inside a ring as prepare_cluster_indexes writes them
    (little-endian, each array padded to a multiple of 8 bytes)

``load_targets`` recognises it by the magic; ``binary_to_text`` / ``text_to_binary`` convert, and the text that
comes back is byte-identical to what prepare_cluster_indexes.py printed (:162-167)."""
import struct

import numpy as np

BINARY_MAGIC = b"WDTGTS\x01\x00"


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


def _pad8(n):
    return (n + 7) // 8 * 8


def save_targets_binary(path, centres, level_offsets, idx, rings):
    """Writes the CSR arrays of a target list (what Engine.ring_query returns) in the binary format above."""
    centres = np.ascontiguousarray(centres, dtype="<u4")
    offs = np.ascontiguousarray(level_offsets, dtype="<u8")
    idx = np.ascontiguousarray(idx, dtype="<u4")
    if offs.size != centres.size * rings + 1 or (offs.size and int(offs[-1]) != idx.size):
        raise ValueError("level_offsets does not describe %d targets x %d rings over %d wells" % (centres.size, rings, idx.size))
    with open(path, "wb") as fh:
        fh.write(BINARY_MAGIC + struct.pack("<IIQQ", rings, 0, centres.size, idx.size))
        for arr in (centres, offs, idx):
            fh.write(arr.tobytes())
            fh.write(b"\0" * (_pad8(arr.nbytes) - arr.nbytes))


class BinaryTargets:
    """A binary target list, memory-mapped: the AllTargets surface count_well_duplicates uses (len, iteration,
    get_all_indices, to_csr) over arrays instead of per-target Python objects."""

    def __init__(self, path, levels=None, limit=None):
        with open(path, "rb") as fh:
            head = fh.read(32)
        if head[:8] != BINARY_MAGIC or len(head) < 32:
            raise ValueError("%s is not a binary target list" % path)
        rings, _, t, n = struct.unpack("<IIQQ", head[8:])
        self._rings_on_file = rings
        pos = 32
        self._centres = np.memmap(path, dtype="<u4", mode="r", offset=pos, shape=(t,))
        pos += _pad8(4 * t)
        self._offs = np.memmap(path, dtype="<u8", mode="r", offset=pos, shape=(t * rings + 1,))
        pos += _pad8(8 * (t * rings + 1))
        self._idx = np.memmap(path, dtype="<u4", mode="r", offset=pos, shape=(n,)) if n else np.zeros(0, "<u4")
        self._t = int(t if not limit else min(t, limit))
        # load_targets(levels=) counts the centre line too and keeps the first `levels` lines of a record
        keep = rings if levels is None else max(0, min(rings, levels - 1))
        self.levels = keep + 1
        if self._t and np.unique(self._centres[:self._t]).size != self._t:
            raise AssertionError("a centre appears twice")          # target.py:72

    def __len__(self):
        return self._t

    def __iter__(self):
        r = self._rings_on_file
        for i in range(self._t):
            o = self._offs[i * r:i * r + self.levels]
            yield Target([[int(self._centres[i])]] + [self._idx[int(o[l]):int(o[l + 1])].tolist() for l in range(self.levels - 1)])

    def to_csr(self, rings=None):
        have = self.levels - 1
        rings = have if rings is None else rings
        if rings > have:
            raise IndexError("list index out of range")
        r = self._rings_on_file
        offs = np.asarray(self._offs[:self._t * r + 1], dtype=np.int64).copy()
        starts = offs[:-1].reshape(self._t, r)[:, :rings] if self._t else np.zeros((0, rings), np.int64)
        ends = offs[1:].reshape(self._t, r)[:, :rings] if self._t else np.zeros((0, rings), np.int64)
        lens = (ends - starts).reshape(-1)
        out_offs = np.zeros(lens.size + 1, np.int64)
        np.cumsum(lens, out=out_offs[1:])
        if rings == r:
            idx = np.asarray(self._idx[:int(offs[-1])], dtype=np.uint32)
        elif rings == 0 or self._t == 0:
            idx = np.zeros(0, np.uint32)
        else:
            # the kept rings of a target are one run of idx (rings 1..rings are stored first)
            idx = np.concatenate([self._idx[int(a):int(b)] for a, b in zip(starts[:, 0], ends[:, rings - 1])]).astype(np.uint32)
        return np.asarray(self._centres[:self._t], dtype=np.uint32), out_offs.astype(np.uint32), idx

    def get_all_indices(self, level=None):
        if level == 0:
            return self._centres[:self._t].tolist()
        centres, offs, idx = self.to_csr()
        if level is None:
            return np.unique(np.concatenate([centres, idx])).tolist()
        rings = self.levels - 1
        sel = np.concatenate([np.arange(offs[i * rings + level - 1], offs[i * rings + level]) for i in range(self._t)]) if self._t else []
        return idx[np.asarray(sel, dtype=np.int64)].tolist()


def write_targets_text(fh, centres, level_offsets, idx, rings, block=4096):
    """The text form (prepare_cluster_indexes.py:162-167): per target the centre, then one comma-joined line per ring."""
    t = len(centres)
    for b0 in range(0, t, block):
        lines = []
        for i in range(b0, min(t, b0 + block)):
            lines.append(str(int(centres[i])))
            for l in range(rings):
                a, b = int(level_offsets[i * rings + l]), int(level_offsets[i * rings + l + 1])
                lines.append(",".join(map(str, idx[a:b].tolist())))
        fh.write("".join(x + "\n" for x in lines))


def binary_to_text(path, fh):
    bt = BinaryTargets(path)
    write_targets_text(fh, bt._centres, bt._offs, bt._idx, bt._rings_on_file)


def text_to_binary(text_path, binary_path):
    at = load_targets(text_path)
    centres, offs, idx = at.to_csr()
    save_targets_binary(binary_path, centres, offs, idx, (at.levels or 1) - 1)


def is_binary_target_file(filename):
    with open(filename, "rb") as fh:
        return fh.read(8) == BINARY_MAGIC


def load_targets(filename, levels=None, limit=None):
    """Reads a target file.  A line without a comma starts a new record; the
    first ``levels`` lines of each record are kept (centre included); reading
    stops after ``limit`` records.  A binary target list (module docstring) is mapped instead of parsed."""
    if is_binary_target_file(filename):
        return BinaryTargets(filename, levels=levels, limit=limit)
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
