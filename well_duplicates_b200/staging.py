"""Host staging pipeline: compressed files of a batch of tiles -> plane buffers the kernels read.

Replaces the per-tile, per-cycle ``gzip.open(...).read()`` of the reference's
reader (bcl_direct_reader.py:200-216 BCL first, CBCL on FileNotFoundError;
:333-345 the .bcl.gz slurp; :292-301 seek + gzip member of a CBCL tile block;
:146-152 the filter header) for the callers that walk whole lanes
(count_well_duplicates.py:207-226).  Gunzip stays on the host
(BASELINE.json north_star) but no longer runs in Python:

* every (tile, cycle) becomes one ``wd_inflate_job`` -- file path, byte range,
  destination -- and ``wd_inflate_batch`` reads and inflates them on native
  threads straight into ONE page-locked block laid out as the kernels want it
  (``[tile][plane][stride]``); the 4-byte cluster count in front of a BCL plane
  lands in the slack before its row, so nothing is copied or re-sliced;
* the block is handed to the library as it is (``wd_tile_map_host``: the
  counting kernel pulls the sectors it needs across PCIe) or, for the dense
  readers (two-pass log mode, exhaustive mode), copied plane by plane by DMA;
* two blocks alternate: the next batch inflates while the GPU counts this one.

The CBCL header and tile table are parsed once per file (CbclFile), not once per
tile and cycle as bcl_direct_reader.py:261-292 does.
"""
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib

LEAD = 256          # slack in front of the first row (takes the count header of plane 0)
BCL_HEADER = 4      # '<I' cluster count in front of the calls (bcl_direct_reader.py:333-338)
PINNED_BUDGET_BYTES = int(os.environ.get("WELLDUP_PINNED_BUDGET", str(3 << 30)))     # per block, two blocks


def _round_up(n, m):
    return (n + m - 1) // m * m


class HostBlock:
    """Grow-only byte block, page-locked (wd_host_alloc) unless ``pinned`` is False
    (CPU-only tests of the host logic)."""

    def __init__(self, pinned=True):
        self.pinned = pinned
        self._pin = None
        self.array = np.empty(0, np.uint8)

    def reserve(self, nbytes):
        if self.array.size >= nbytes:
            return
        if self.pinned:
            from .engine import PinnedArray
            if self._pin is not None:
                self._pin.free()
            self._pin = PinnedArray((nbytes,))
            self.array = self._pin.array
        else:
            self.array = np.empty(nbytes, np.uint8)

    def free(self):
        if self._pin is not None:
            self._pin.free()
            self._pin = None
        self.array = np.empty(0, np.uint8)


class StagedBatch:
    """Inflated planes and filters of a batch of same-sized tiles, in one HostBlock."""

    def __init__(self, tiles, cycles, block, n_clusters, stride, fstride):
        self.tiles = tiles
        self.cycles = cycles                    # sorted unique 0-based cycles = plane order in the block
        self.plane_of = {c: p for p, c in enumerate(cycles)}
        self.block = block
        self.n_clusters = n_clusters
        self.stride = stride
        self.fstride = fstride
        n, p = len(tiles), len(cycles)
        self.kinds = np.full((n, p), _lib.PLANE_BCL, np.uint8)
        self.n_block = np.full((n, p), n_clusters, np.uint32)
        self.usize = np.full((n, p), n_clusters, np.uint32)
        self.compressed_bytes = 0
        self.inflated_bytes = 0

    def planes(self, k):
        """[planes, stride] view of tile k."""
        p = len(self.cycles)
        off = LEAD + k * p * self.stride
        return self.block.array[off:off + p * self.stride].reshape(p, self.stride)

    def filter(self, k):
        off = LEAD + len(self.tiles) * len(self.cycles) * self.stride + k * self.fstride
        return self.block.array[off:off + self.n_clusters]


class Stager:
    def __init__(self, pinned=True, threads=None, cbcl_cache=None):
        self._lib = _lib.load()
        self.blocks = [HostBlock(pinned), HostBlock(pinned)]
        # default: the cores this process may run on (affinity mask / cgroup), not every core of the box
        usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.threads = int(threads if threads else int(os.environ.get("WELLDUP_INFLATE_THREADS", "0")) or usable)
        self._cbcl_cache = {} if cbcl_cache is None else cbcl_cache
        self._worker = ThreadPoolExecutor(max_workers=1)

    def close(self):
        self._worker.shutdown(wait=True)
        for b in self.blocks:
            b.free()

    def tiles_per_batch(self, n_clusters, n_cycles, limit=4096):
        per_tile = max(1, n_cycles) * _round_up(n_clusters + BCL_HEADER, 256) + _round_up(n_clusters, 256)
        return int(max(1, min(limit, PINNED_BUDGET_BYTES // per_tile)))

    def reserve(self, which, n_clusters, n_tiles, n_planes):
        """Room in block ``which`` for a batch.  Page-locked memory is allocated here, on the
        caller's thread (the one that owns the CUDA device), not on the inflate worker."""
        stride = _round_up(n_clusters + BCL_HEADER, 256)
        self.blocks[which].reserve(LEAD + n_tiles * (n_planes * stride + _round_up(n_clusters, 256)))

    # ---- files -> block (no CUDA calls: safe on the worker thread) --------------------------------
    def _run(self, jobs):
        """wd_inflate_batch over a list of (path, offset, size, dst_address, cap)."""
        arr = (_lib.InflateJob * max(1, len(jobs)))()
        keep = [os.fsencode(path) for path, _, _, _, _ in jobs]      # alive for the duration of the call
        for j, enc, (_, off, size, dst, cap) in zip(arr, keep, jobs):
            j.path = enc
            j.offset = off
            j.size = size
            j.dst = dst
            j.dst_cap = cap
        if jobs:
            self._lib.wd_inflate_batch(arr, len(jobs), self.threads)     # per-job statuses are read by the caller
        return arr

    def load(self, tiles, cycles, which=0):
        """Filters and the planes of ``cycles`` (0-based) of ``tiles`` (reader.Tile objects with
        one cluster count) -> StagedBatch in block ``which``."""
        from .reader import CbclFile
        uniq = sorted(set(cycles))
        n = tiles[0].num_clusters
        if not all(t.num_clusters == n for t in tiles):
            raise ValueError("the tiles of one batch must have the same cluster count")
        stride = _round_up(n + BCL_HEADER, 256)
        fstride = _round_up(n, 256)
        self.reserve(which, n, len(tiles), len(uniq))        # no-op when lane_batches made room already
        block = self.blocks[which]
        batch = StagedBatch(tiles, uniq, block, n, stride, fstride)
        base = block.array.ctypes.data
        # filters: plain files
        for k, t in enumerate(tiles):
            view = block.array[LEAD + len(tiles) * len(uniq) * stride + k * fstride:][:fstride]
            with open(t.filter_file, "rb") as fh:
                if struct.unpack("<III", fh.read(12)) != (0, 3, n):          # bcl_direct_reader.py:146-152
                    raise AssertionError("%s: filter header is not (0, 3, %d)" % (t.filter_file, n))
                got = fh.readinto(memoryview(view[:n]))
            view[got:] = 0
        # planes: BCL first ...
        jobs, where = [], []
        for k, t in enumerate(tiles):
            for p, cyc in enumerate(uniq):
                row = base + LEAD + (k * len(uniq) + p) * stride
                jobs.append((os.path.join(t.data_dir, "C%i.1" % (cyc + 1), t.bcl_filename), 0, 0,
                             row - BCL_HEADER, n + BCL_HEADER))
                where.append((k, p, cyc, row))
        res = self._run(jobs)
        # ... CBCL where there is no BCL file (bcl_direct_reader.py:209-216)
        cjobs, cwhere = [], []
        for j, (k, p, cyc, row) in zip(res, where):
            if j.status != _lib.WD_E_NOENT:
                continue
            t = tiles[k]
            path = os.path.join(t.data_dir, "C%i.1" % (cyc + 1), t.cbcl_filename)
            cf = self._cbcl_cache.get(path)
            if cf is None:
                cf = self._cbcl_cache[path] = CbclFile(path)
            if int(t.tile) not in cf.blocks:
                raise AssertionError("%s holds no block for tile %s" % (path, t.tile))
            off, ncl, usize, csize = cf.blocks[int(t.tile)]
            if usize > stride:
                raise AssertionError("CBCL block of %d bytes for a tile of %d clusters" % (usize, n))
            cjobs.append((path, off, csize, row, usize))
            cwhere.append((k, p, ncl, usize, cf.excluded))
        cres = self._run(cjobs)
        cres_of = {(k, p): (j, ncl, usize, excl) for j, (k, p, ncl, usize, excl) in zip(cres, cwhere)}
        # verdicts, in the order the reference would meet the files
        for j, (k, p, cyc, row) in zip(res, where):
            if j.status == _lib.WD_E_NOENT:
                cj, ncl, usize, excl = cres_of[(k, p)]
                # GzipFile.read(usize) stops after usize bytes whatever follows (bcl_direct_reader.py:301)
                if cj.status not in (_lib.WD_OK, _lib.WD_E_CAPACITY):
                    _lib.raise_status(cj.status, cj.message.decode("utf-8", "replace"))
                if cj.out_len * 2 < ncl:
                    raise AssertionError("CBCL block of %d bytes cannot hold %d clusters" % (cj.out_len, ncl))
                batch.kinds[k, p] = _lib.PLANE_CBCL_EXCL if excl else _lib.PLANE_CBCL
                batch.n_block[k, p] = ncl
                batch.usize[k, p] = cj.out_len
                batch.compressed_bytes += int(cj.size)
                batch.inflated_bytes += int(cj.out_len)
                continue
            if j.status not in (_lib.WD_OK, _lib.WD_E_CAPACITY):
                _lib.raise_status(j.status, j.message.decode("utf-8", "replace"))
            off = row - BCL_HEADER - base
            (count,) = struct.unpack("<I", block.array[off:off + BCL_HEADER].tobytes()) if j.out_len >= BCL_HEADER else (None,)
            if count != n:                                       # bcl_direct_reader.py:338
                raise AssertionError("BCL header says %s clusters, filter says %d" % (count, n))
            # stricter than the reference, which fails only when a requested well lies beyond the data it got:
            # a plane that is not exactly one call per well is refused (intentional divergences)
            if j.out_len != n + BCL_HEADER:
                raise AssertionError("BCL file holds %d calls, header says %d" % (j.out_len - BCL_HEADER, n))
            batch.inflated_bytes += int(j.out_len)
        return batch

    def load_async(self, tiles, cycles, which=0):
        return self._worker.submit(self.load, tiles, cycles, which)

    # ---- block -> engine (caller's thread) -------------------------------------------------------
    def deliver(self, engine, batch, first_slot=0, zero_copy=True):
        """Make tile k of the batch tile slot ``first_slot + k``; returns {cycle: plane index}."""
        for k in range(len(batch.tiles)):
            slot = first_slot + k
            if zero_copy:
                engine.tile_map_host(slot, batch.n_clusters, batch.planes(k), kinds=batch.kinds[k],
                                     n_block=batch.n_block[k], pinned_filter=batch.filter(k))
                continue
            engine.tile_begin(slot, batch.n_clusters, len(batch.cycles))
            engine.tile_put_filter(slot, batch.filter(k))
            planes = batch.planes(k)
            for p in range(len(batch.cycles)):
                if batch.kinds[k, p] == _lib.PLANE_BCL:
                    engine.tile_put_bcl(slot, p, planes[p, :batch.n_clusters])
                else:
                    engine.tile_put_cbcl(slot, p, planes[p, :batch.usize[k, p]], int(batch.n_block[k, p]),
                                         batch.kinds[k, p] == _lib.PLANE_CBCL_EXCL)
        return dict(batch.plane_of)


def lane_batches(stager, open_tile, names, cycles, per_batch=None, announce=None, on_error=None):
    """Walks ``names`` (tile names, in order): yields (names of the batch, StagedBatch)
    while the following batch is read and inflated in the background.
    ``open_tile(name)`` -> reader.Tile.  A batch holds tiles of one cluster count, at
    most ``per_batch`` of them (default: what the pinned budget allows).
    ``announce(name)`` is called for each tile just before its batch is waited for --
    with per_batch=1 that is the moment the reference logs "Reading tile".
    Errors surface where the reference would meet them (count_well_duplicates.py:207-226
    walks tile by tile): every tile in front of the failing one is handed out first -- a
    tile that cannot be opened ends its batch early and the error is kept for the next
    round; a batch whose files fail to load is loaded again tile by tile -- and
    ``on_error(name)`` is called with the failing tile just before the exception is raised."""
    uniq = sorted(set(cycles))
    pending = {}                           # index of a tile that failed to open -> its exception

    def prepare(start, which, limit=None):
        """Caller's thread: open the tiles of the next batch, make room for them (page-locked
        memory is allocated on the thread that owns the CUDA device), start the inflate."""
        tiles = []
        if start in pending:
            return tiles, None, pending.pop(start)
        limit = per_batch if limit is None else limit
        k = start
        try:
            while k < len(names):
                t = open_tile(names[k])
                if tiles and t.num_clusters != tiles[0].num_clusters:
                    break
                if limit is None:
                    limit = stager.tiles_per_batch(t.num_clusters, len(uniq))
                tiles.append(t)
                k += 1
                if len(tiles) >= limit:
                    break
            if tiles:
                stager.reserve(which, tiles[0].num_clusters, len(tiles), len(uniq))
        except Exception as exc:           # noqa: BLE001 -- raised when the failing tile is due
            if not tiles:
                return tiles, None, exc
            pending[k] = exc               # the tiles opened so far are a batch of their own
            try:
                stager.reserve(which, tiles[0].num_clusters, len(tiles), len(uniq))
            except Exception as exc2:      # noqa: BLE001
                return [], None, exc2
        return tiles, (stager.load_async(tiles, uniq, which) if tiles else None), None

    def fail(name, exc):
        if on_error is not None:
            on_error(name)
        raise exc

    start, which = 0, 0
    tiles, fut, err = prepare(start, which)
    while start < len(names):
        if announce is not None and per_batch == 1:
            announce(names[start])
        if err is not None:
            fail(names[start], err)
        try:
            batch = fut.result()
        except Exception as exc:           # noqa: BLE001
            if len(tiles) == 1:
                fail(names[start], exc)
            # some file of the batch is bad: tile by tile, so that the good ones in front of it are counted
            for k in range(start, start + len(tiles)):
                one, f1, e1 = prepare(k, which, limit=1)
                if e1 is not None:
                    fail(names[k], e1)
                try:
                    b1 = f1.result()
                except Exception as exc1:  # noqa: BLE001
                    fail(names[k], exc1)
                yield names[k:k + 1], b1
            start += len(tiles)            # (every tile loaded on its own: carry on)
            tiles, fut, err = prepare(start, which)
            continue
        got = names[start:start + len(tiles)]
        start += len(tiles)
        which ^= 1
        tiles, fut, err = prepare(start, which)
        yield got, batch
