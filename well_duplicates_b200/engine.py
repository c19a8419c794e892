"""Thin object wrapper over the C ABI: one ``Engine`` per GPU.

Nothing here computes on the CPU: arrays go in, the CUDA kernels run, arrays
come out.  Creating an Engine without a usable B200 raises."""
import ctypes as C
import weakref

import numpy as np

from . import _lib

WINDOW_LO = 20000      # prepare_cluster_indexes.py:43,52
WINDOW_HI = 20001      # ... the break is tested after the append (:61-67)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def bind_to_gpu_numa_node(device):
    """Multi-GPU boxes: run this process on the cores next to its GPU, so that the pinned staging memory
    it allocates (first touch) sits on the NUMA node the GPU's PCIe link hangs off.  Eight ranks whose
    staging memory ends up on one socket halve each other's host-to-device bandwidth.  Returns the PCI
    address and the cores chosen, or None when the platform does not say (then nothing is changed)."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(int(device))],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        bdf = out[0].strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                    # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return bdf, sorted(cpus)
    except Exception:
        return None


class PinnedArray:
    """numpy view over page-locked host memory from wd_host_alloc."""

    def __init__(self, shape, dtype=np.uint8):
        lib = _lib.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _lib.check(lib.wd_host_alloc(max(self.nbytes, 1), C.byref(p)))
        self._p = p
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._fin = weakref.finalize(self, lib.wd_host_free, p)

    def free(self):
        self.array = None
        self._fin()


class RegisteredArray:
    """Page-locks a numpy array the caller already has (wd_host_register), so that it can be given to
    Engine.tile_map_host like a PinnedArray; ``release()`` (or garbage collection) unlocks it."""

    def __init__(self, array):
        if not array.flags.c_contiguous:
            raise ValueError("only a C-contiguous array can be page-locked in place")
        lib = _lib.load()
        self.array = array
        _lib.check(lib.wd_host_register(_ptr(array), array.nbytes))
        self._fin = weakref.finalize(self, lib.wd_host_unregister, C.c_void_p(array.ctypes.data))

    def release(self):
        self._fin()


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.wd_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, self._lib.wd_destroy, h)
        self.n_targets = 0
        self.levels = 0
        self._slots = {}

    def close(self):
        self._fin()

    # ---- plumbing ----------------------------------------------------------
    def set_stream(self, cuda_stream):
        _lib.check(self._lib.wd_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def sync(self):
        _lib.check(self._lib.wd_sync(self._h))

    def set_l2_fetch_granularity(self, nbytes):
        """-> previous value of cudaLimitMaxL2FetchGranularity."""
        prev = C.c_int32()
        _lib.check(self._lib.wd_set_l2_fetch_granularity(self._h, int(nbytes), C.byref(prev)))
        return prev.value

    def set_tuning(self, step0=0, step1=0, centre_chunk=0, head_planes=-1, head_groups=0, visit_order=-1, targets_per_cta=0,
                   ctas_per_sm=0):
        """Measurement knobs of wd_count (struct wd_tuning); no arguments = the library's defaults."""
        t = _lib.Tuning(step0=int(step0), step1=int(step1), centre_chunk=int(centre_chunk), head_planes=int(head_planes),
                        head_groups=int(head_groups), visit_order=int(visit_order),
                        targets_per_cta=int(targets_per_cta), ctas_per_sm=int(ctas_per_sm))
        _lib.check(self._lib.wd_set_tuning(self._h, C.byref(t)))

    def launch_count(self):
        n = C.c_uint64()
        _lib.check(self._lib.wd_launch_count(self._h, C.byref(n)))
        return n.value

    def last_count_h2d_bytes(self):
        n = C.c_uint64()
        _lib.check(self._lib.wd_last_count_h2d_bytes(self._h, C.byref(n)))
        return n.value

    def last_count_staging(self):
        """-> (head planes the last count copied by DMA, GB/s measured for an earlier count's copies or 0.0)."""
        h, r = C.c_int32(), C.c_double()
        _lib.check(self._lib.wd_last_count_staging(self._h, C.byref(h), C.byref(r)))
        return h.value, r.value

    # ---- stage 1 --------------------------------------------------------------
    def load_locs(self, xy):
        xy = _c(xy, np.float32).reshape(-1, 2)
        self.n_locs = xy.shape[0]
        _lib.check(self._lib.wd_locs_load(self._h, _ptr(xy), xy.shape[0]))

    def pixels(self):
        x = np.empty(self.n_locs, np.int32)
        y = np.empty(self.n_locs, np.int32)
        _lib.check(self._lib.wd_locs_pixels(self._h, _ptr(x), _ptr(y)))
        return x, y

    def ring_query(self, centres, levels=5, window_lo=WINDOW_LO, window_hi=WINDOW_HI):
        """-> (level_offsets[t*levels+1], idx) as uint32 arrays."""
        centres = _c(centres, np.uint32)
        t = centres.size
        offs = np.zeros(t * levels + 1, np.uint32)
        cap = max(1024, t * 32 * levels)
        n_idx = C.c_uint64()
        first_empty = C.c_uint32()
        while True:
            idx = np.empty(cap, np.uint32)
            rc = self._lib.wd_ring_query(self._h, _ptr(centres), t, levels, window_lo, window_hi, _ptr(offs),
                                         _ptr(idx), cap, C.byref(n_idx), C.byref(first_empty))
            if rc == _lib.WD_E_CAPACITY:
                cap = int(n_idx.value)
                continue
            if rc == _lib.WD_E_RUNTIME:
                err = RuntimeError(self._lib.wd_last_error().decode())
                err.first_empty = (first_empty.value // levels, first_empty.value % levels)
                raise err
            _lib.check(rc)
            return offs, idx[: n_idx.value].copy()

    # ---- target list ------------------------------------------------------------
    def load_targets(self, centres, level_offsets, idx, levels):
        centres = _c(centres, np.uint32)
        level_offsets = _c(level_offsets, np.uint32)
        idx = _c(idx, np.uint32)
        if idx.size == 0:
            idx = np.zeros(1, np.uint32)
        _lib.check(self._lib.wd_targets_load(self._h, _ptr(centres), _ptr(level_offsets), _ptr(idx),
                                             centres.size, int(levels)))
        self.n_targets = int(centres.size)
        self.levels = int(levels)

    # ---- tile staging -------------------------------------------------------------
    def tile_begin(self, slot, n_clusters, n_planes):
        _lib.check(self._lib.wd_tile_begin(self._h, slot, n_clusters, n_planes))
        self._slots[slot] = (int(n_clusters), int(n_planes))

    def tile_put_filter(self, slot, filt):
        filt = _c(filt, np.uint8)
        _lib.check(self._lib.wd_tile_put_filter(self._h, slot, _ptr(filt), filt.size))

    def tile_put_bcl(self, slot, plane, data):
        data = _c(data, np.uint8)
        _lib.check(self._lib.wd_tile_put_bcl(self._h, slot, plane, _ptr(data), data.size))

    def tile_put_cbcl(self, slot, plane, nibbles, n_block, excluded):
        nibbles = _c(nibbles, np.uint8)
        _lib.check(self._lib.wd_tile_put_cbcl(self._h, slot, plane, _ptr(nibbles), nibbles.size, int(n_block),
                                              1 if excluded else 0))

    def tile_map_host(self, slot, n_clusters, pinned_planes, kinds=None, n_block=None, pinned_filter=None):
        """Zero-copy staging: `pinned_planes` is a 2-D uint8 view [n_planes, stride] of a
        PinnedArray (`pinned_filter` optionally the filter bytes, likewise pinned); the
        kernels read them in place across PCIe."""
        a = pinned_planes
        if a.dtype != np.uint8 or a.ndim != 2 or not a.flags.c_contiguous:
            raise ValueError("tile_map_host wants a C-contiguous 2-D uint8 array in pinned memory")
        kinds = None if kinds is None else _c(kinds, np.uint8)
        n_block = None if n_block is None else _c(n_block, np.uint32)
        _lib.check(self._lib.wd_tile_map_host(self._h, slot, int(n_clusters), a.shape[0], _ptr(a), a.shape[1],
                                              None if kinds is None else _ptr(kinds),
                                              None if n_block is None else _ptr(n_block),
                                              None if pinned_filter is None else _ptr(pinned_filter)))
        self._slots[slot] = (int(n_clusters), int(a.shape[0]))

    def filter_offsets(self, slot):
        n = self._slots[slot][0]
        out = np.empty(n, np.int32)
        passing = C.c_uint32()
        _lib.check(self._lib.wd_filter_offsets(self._h, slot, _ptr(out), C.byref(passing)))
        return out, passing.value

    def get_seqs(self, slot, indices, plane_order):
        """-> (codes uint8 [n, len] with 0..3 = ACGT, 4 = N; pf uint8 [n])."""
        indices = _c(indices, np.int64)
        order = _c(plane_order, np.int32)
        codes = np.empty((indices.size, order.size), np.uint8)
        pf = np.empty(indices.size, np.uint8)
        _lib.check(self._lib.wd_get_seqs(self._h, slot, _ptr(indices), indices.size, _ptr(order), order.size,
                                         _ptr(codes), _ptr(pf)))
        return codes, pf

    # ---- stage 3 ---------------------------------------------------------------------
    def count(self, first_slot, n_tiles, plane_order, edit_distance=2, hamming=False, mode=0, per_target=True):
        """-> (per_target int32 [tiles, T, 1+2L] or None, counters int64 [tiles, 1+5L])."""
        order = _c(plane_order, np.int32)
        L = self.levels
        counters = np.empty((n_tiles, 1 + 5 * L), np.int64)
        pt = np.empty((n_tiles, self.n_targets, 1 + 2 * L), np.int32) if per_target else None
        _lib.check(self._lib.wd_count(self._h, first_slot, n_tiles, _ptr(order), order.size, int(edit_distance),
                                      1 if hamming else 0, int(mode), _ptr(pt) if per_target else None,
                                      _ptr(counters)))
        self._last = (n_tiles, per_target)
        self._last_len = int(order.size)
        return pt, counters

    def count_async(self, first_slot, n_tiles, plane_order, edit_distance=2, hamming=False, mode=0,
                    per_target=False):
        order = _c(plane_order, np.int32)
        _lib.check(self._lib.wd_count_async(self._h, first_slot, n_tiles, _ptr(order), order.size,
                                            int(edit_distance), 1 if hamming else 0, int(mode),
                                            1 if per_target else 0))
        self._last = (n_tiles, per_target)
        self._last_len = int(order.size)

    def trace_sectors(self, first_slot, n_tiles, plane_order, edit_distance=2, hamming=False):
        """Measurement hook (wd_count_trace_sectors): -> (sectors, lines) uint32 [n_tiles, len], the distinct
        32-byte sectors / 128-byte lines of each compared position's plane the fused kernel reads."""
        order = _c(plane_order, np.int32)
        sectors = np.zeros((n_tiles, order.size), np.uint32)
        lines = np.zeros((n_tiles, order.size), np.uint32)
        _lib.check(self._lib.wd_count_trace_sectors(self._h, first_slot, n_tiles, _ptr(order), order.size,
                                                    int(edit_distance), 1 if hamming else 0, _ptr(sectors), _ptr(lines)))
        return sectors, lines

    def count_fetch(self):
        n_tiles, per_target = self._last
        L = self.levels
        counters = np.empty((n_tiles, 1 + 5 * L), np.int64)
        pt = np.empty((n_tiles, self.n_targets, 1 + 2 * L), np.int32) if per_target else None
        _lib.check(self._lib.wd_count_fetch(self._h, _ptr(pt) if per_target else None, _ptr(counters)))
        return pt, counters

    def dup_pairs(self, with_seqs=False):
        """Rows (tile of the batch, target ordinal, well, distance) of the last count in mode 1 or 2, in the
        order the reference logs them; with_seqs: also codes uint8 [n, 2, len] (centre, well; 0..3 ACGT, 4 N)."""
        n = C.c_uint64()
        rc = self._lib.wd_dup_pairs_seqs(self._h, None, None, 0, C.byref(n))
        if rc not in (_lib.WD_OK, _lib.WD_E_CAPACITY):
            _lib.check(rc)
        rows = np.empty((max(int(n.value), 1), 4), np.int32)
        codes = np.empty((rows.shape[0], 2, self._last_len), np.uint8) if with_seqs else None
        _lib.check(self._lib.wd_dup_pairs_seqs(self._h, _ptr(rows), _ptr(codes) if with_seqs else None, rows.shape[0],
                                               C.byref(n)))
        if with_seqs:
            return rows[: n.value], codes[: n.value]
        return rows[: n.value]

    def publish_counters(self, tile_row, lane_row, n_rows_total, add=False):
        """K7: -> (device pointer, n_int64) of the zero-padded all-reduce buffer; ``add``: into the
        buffer of the previous call (a rank that counts its tiles in several batches)."""
        tile_row = _c(tile_row, np.int32)
        lane_row = _c(lane_row, np.int32)
        p = C.c_void_p()
        n = C.c_size_t()
        fn = self._lib.wd_publish_add if add else self._lib.wd_publish_counters
        _lib.check(fn(self._h, _ptr(tile_row), _ptr(lane_row), tile_row.size, int(n_rows_total), C.byref(p), C.byref(n)))
        return p.value, n.value

    def comm_join(self):
        """The engine's stream waits for the all-reduces issued so far."""
        _lib.check(self._lib.wd_comm_join(self._h))

    # ---- multi-GPU (NCCL inside the library, no torch) -------------------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes that rank 0 makes and every rank passes to comm_init."""
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        _lib.check(_lib.load().wd_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, nranks):
        _lib.check(self._lib.wd_comm_init(self._h, C.c_char_p(bytes(unique_id)), int(rank), int(nranks)))

    def comm_destroy(self):
        _lib.check(self._lib.wd_comm_destroy(self._h))

    def allreduce_published(self):
        """ncclAllReduce(int64, sum), in place, of the buffer of the last publish_counters."""
        _lib.check(self._lib.wd_allreduce_i64(self._h, None, 0))

    def published_fetch(self, n_int64):
        out = np.empty(int(n_int64), np.int64)
        _lib.check(self._lib.wd_published_fetch(self._h, _ptr(out), out.size))
        return out

    def count_exhaustive(self, slot, plane_order, levels=5, edit_distance=2, hamming=False,
                         window_lo=WINDOW_LO, window_hi=WINDOW_HI):
        order = _c(plane_order, np.int32)
        counters = np.empty(1 + 5 * levels, np.int64)
        _lib.check(self._lib.wd_count_exhaustive(self._h, slot, _ptr(order), order.size, int(levels), window_lo,
                                                 window_hi, int(edit_distance), 1 if hamming else 0,
                                                 _ptr(counters)))
        return counters
