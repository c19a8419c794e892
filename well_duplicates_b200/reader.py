"""Reads base calls of selected wells straight from a run folder.

Same interface as the reference's bcl_direct_reader.py -- ``BCLReader(location)
.get_tile(lane, tile).get_seqs(indices, start, end) -> {idx: (str, bool)}`` with
``SEQUENCE = 0`` / ``QUAL_FLAG = 1`` -- but the host only finds files and
inflates them; the per-well gather, the BCL / CBCL decode and the filter-offset
table run on the GPU (K3, K4, K5 behind wd_get_seqs / wd_filter_offsets).

File layout handled (reference lines in brackets):
  L00x/<prefix>_<tile>.filter           [bcl_direct_reader.py:123-152]
  L00x/C<n>.1/<prefix>_<tile>.bcl.gz    [:200-210, :327-345]
  L00x/C<n>.1/L00x_<surface>.cbcl       [:137, :255-301; cbcl_read.py:20-84]
"""
import os
import re
import struct

import numpy as np

from . import _lib
from .engine import Engine

SEQUENCE = 0
QUAL_FLAG = 1

_LUT = np.frombuffer(b"ACGTN", dtype=np.uint8)
_engine = None
_stager = None


def default_engine():
    """Process-wide Engine on the GPU named by LOCAL_RANK (0 if unset)."""
    global _engine
    if _engine is None:
        _engine = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _engine


def default_stager(cbcl_cache=None):
    """Process-wide staging.Stager (two page-locked blocks, native inflate threads)."""
    global _stager
    if _stager is None:
        from .staging import Stager
        _stager = Stager()
    if cbcl_cache is not None:
        _stager._cbcl_cache = cbcl_cache
    return _stager


def gunzip(data, size_hint=None):
    """All members of a gzip byte string (gzip.open(...).read() semantics), inflated by
    the library (wd_gunzip).  ``size_hint`` = expected inflated size, if known."""
    import ctypes as C
    lib = _lib.load()
    data = bytes(data)
    cap = int(size_hint) if size_hint else max(4 * len(data), 1 << 16)
    while True:
        out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        rc = lib.wd_gunzip(data, len(data), out.ctypes.data, cap, C.byref(n))
        if rc != _lib.WD_E_CAPACITY:
            _lib.check(rc)
            return out[:n.value].tobytes()
        cap *= 4


def codes_to_strings(codes):
    chars = _LUT[codes]
    return [row.tobytes().decode("ascii") for row in chars]


class CbclFile:
    """Header + tile table of one CBCL file, parsed once and kept."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as fh:
            version, hsize, bbits, qbits, nbins = struct.unpack("<HIBBI", fh.read(12))
            assert version == 1
            assert hsize > 32
            assert bbits == 2
            assert qbits == 2
            assert nbins == 4
            rest = fh.read(nbins * 8 + 4)
            (tcount,) = struct.unpack("<I", rest[-4:])
            table = fh.read(tcount * 16 + 1)
        self.excluded = bool(table[-1])
        rec = np.frombuffer(table[:-1], dtype="<u4").reshape(tcount, 4)
        starts = hsize + np.concatenate([[0], np.cumsum(rec[:, 3].astype(np.int64))[:-1]])
        self.blocks = {}
        for (tno, ncl, usize, csize), off in zip(rec.tolist(), starts.tolist()):
            self.blocks.setdefault(tno, (off, ncl, usize, csize))   # first match wins, like the reference's loop

    def read_tile(self, tile):
        """-> (inflated block as uint8 array, cluster count in the record, excluded flag)."""
        assert int(tile) in self.blocks
        off, ncl, usize, csize = self.blocks[int(tile)]
        with open(self.path, "rb") as fh:
            fh.seek(off)
            comp = fh.read(csize)
        import ctypes as C
        out = np.empty(max(usize, 1), np.uint8)
        n = C.c_size_t()
        rc = _lib.load().wd_gunzip(comp, len(comp), out.ctypes.data, usize, C.byref(n))
        if rc != _lib.WD_E_CAPACITY:            # GzipFile.read(usize) stops after usize bytes (:301)
            _lib.check(rc)
        data = out[:n.value]
        return data, ncl, self.excluded


class BCLReader(object):
    def __init__(self, location=".", engine=None):
        base = os.listdir(os.path.join(location, "Data", "Intensities", "BaseCalls"))
        self.lanes = [d for d in base if re.match(r"L\d\d\d$", d)]
        self.location = location
        self._engine = engine
        self._cbcl_cache = {}

    @property
    def engine(self):
        return self._engine if self._engine is not None else default_engine()

    def get_seq(self, lane, tile, cluster_index, start=0, end=None):
        return self.get_tile(lane, tile).get_seqs([cluster_index], start, end)[cluster_index]

    def get_tile(self, lane, tile, in_memory=False):
        lane_dir = str(lane)
        if lane_dir not in self.lanes:
            lane_dir = "L%03d" % int(lane_dir)
        if in_memory:
            raise RuntimeError("Preloading into memory not implemented yet")
        return Tile(os.path.join(self.location, "Data", "Intensities", "BaseCalls", lane_dir), tile,
                    engine=self._engine, cbcl_cache=self._cbcl_cache)


class Tile(object):
    def __init__(self, data_dir, tile, engine=None, cbcl_cache=None):
        self.data_dir = data_dir
        self.tile = tile
        self._engine = engine
        self._cbcl_cache = {} if cbcl_cache is None else cbcl_cache
        self.bcl_filename = None
        listing = os.listdir(data_dir)
        for name in listing:
            m = re.match("(.+_%s).filter" % tile, name)
            if m:
                self.bcl_filename = m.group(1) + ".bcl.gz"
                self.filter_file = os.path.join(data_dir, name)
                break
        if not self.bcl_filename:
            raise RuntimeError("Cannot find a .filter file for tile %s" % tile)
        self.cbcl_filename = "%s_%s.cbcl" % (os.path.basename(data_dir), str(tile)[0])
        self.num_cycles = len([f for f in listing if re.match(r"C\d+.1$", f)])
        with open(self.filter_file, "rb") as fh:
            hdr = struct.unpack("<III", fh.read(12))
            assert tuple(hdr[0:2]) == (0, 3)
            self.num_clusters = hdr[2]
        self.filter_offsets = None
        self.passing_wells = None

    @property
    def engine(self):
        return self._engine if self._engine is not None else default_engine()

    # ---- host side: files -> bytes -----------------------------------------------
    def read_filter(self):
        with open(self.filter_file, "rb") as fh:
            assert struct.unpack("<III", fh.read(12)) == (0, 3, self.num_clusters)
            return np.frombuffer(fh.read(), dtype=np.uint8)[: self.num_clusters]

    def stage(self, slot, cycles, pool=None, zero_copy=False):
        """Load the filter and the planes of ``cycles`` (0-based, any order, may
        repeat) into tile slot ``slot``; returns {cycle: plane index}.  Files are read
        and inflated by the library's native threads (staging.Stager) into page-locked
        memory; ``zero_copy`` leaves them there for the kernels to read in place (the
        block is reused by the next stage() call), otherwise they are copied to HBM."""
        eng = self.engine
        st = default_stager(self._cbcl_cache)
        batch = st.load([self], cycles, which=0)
        zero_copy = zero_copy and len(batch.cycles) > 0
        plane_of = st.deliver(eng, batch, first_slot=slot, zero_copy=zero_copy)
        if not zero_copy:
            eng.sync()          # the copies read the block the next call overwrites
        return plane_of

    # ---- reference API -------------------------------------------------------------
    def get_seqs(self, cluster_indices, start=0, end=None):
        if end is None:
            end = self.num_cycles
        keys = sorted({int(i) for i in cluster_indices})
        if keys[-1] >= self.num_clusters:
            raise IndexError("Requested cluster %i is out of range.  Highest on this tile is %i." %
                             (keys[-1], self.num_clusters - 1))
        if keys[0] < 0:
            raise IndexError("Requested cluster %i is a negative number." % keys[0])
        cycles = list(range(start, end))
        # Few wells: the gather pulls their sectors across PCIe (~0.3 G requests/s) instead of copying
        # whole planes (~50 GB/s); the two cost the same when about 1 well in 170 is asked for.
        plane_of = self.stage(0, cycles, zero_copy=len(keys) * 128 < self.num_clusters)
        codes, pf = self.engine.get_seqs(0, keys, [plane_of[c] for c in cycles])
        seqs = codes_to_strings(codes) if cycles else [""] * len(keys)
        return {k: (s, bool(f)) for k, s, f in zip(keys, seqs, pf)}

    def _get_filter_offsets(self):
        if self.filter_offsets:
            return self.filter_offsets
        eng = self.engine
        eng.tile_begin(0, self.num_clusters, 0)
        eng.tile_put_filter(0, self.read_filter())
        offs, passing = eng.filter_offsets(0)
        self.filter_offsets = offs.tolist()
        self.passing_wells = passing
        return self.filter_offsets
