"""ctypes wrapper of oracle/welldup_oracle.c -- TEST INFRASTRUCTURE ONLY (see the
header of that file).  Used where the pure-Python restatement (ref_port.py) is
too slow: medium / full-size parity checks and the CPU baseline of bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None
_u8p = C.POINTER(C.c_uint8)


def build():
    src = os.path.join(HERE, "welldup_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_filter_offsets.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def locs_to_pixels(xy):
    xy = np.ascontiguousarray(xy, np.float32)
    n = xy.shape[0]
    x, y = np.empty(n, np.int32), np.empty(n, np.int32)
    lib().orc_locs_to_pixels(_p(xy), C.c_uint32(n), _p(x), _p(y))
    return x, y


def ring_indexes(X, Y, centre, levels=5, cap=4096):
    X = np.ascontiguousarray(X, np.int32)
    Y = np.ascontiguousarray(Y, np.int32)
    out = np.empty(levels * cap, np.uint32)
    counts = np.zeros(levels, np.uint32)
    rc = lib().orc_ring_indexes(_p(X), _p(Y), C.c_uint32(X.size), C.c_uint32(centre), levels, _p(out), _p(counts),
                                C.c_uint32(cap))
    if rc == -100:
        return ring_indexes(X, Y, centre, levels, cap * 8)
    if rc < 0:
        raise RuntimeError("Got no wells for cluster %s at (%s,%s) level %s", (centre, int(X[centre]), int(Y[centre]), -1 - rc))
    return [out[l * cap: l * cap + counts[l]].tolist() for l in range(levels)]


def rings_csr(X, Y, centres, levels=5):
    """-> (level_offsets, idx) uint32 arrays for many centres."""
    offs = [0]
    idx = []
    for c in centres:
        for ring in ring_indexes(X, Y, int(c), levels):
            idx.extend(ring)
            offs.append(len(idx))
    return np.array(offs, np.uint32), np.array(idx, np.uint32)


def filter_offsets(filt):
    filt = np.ascontiguousarray(filt, np.uint8)
    off = np.empty(filt.size, np.int32)
    passing = lib().orc_filter_offsets(_p(filt), C.c_uint32(filt.size), _p(off))
    return off, int(passing)


KIND = {"bcl": 1, "cbcl": 2, "cbcl_excl": 3}


def _plane_ptrs(planes):
    keep = [np.ascontiguousarray(p, np.uint8) for p in planes]
    arr = (C.c_void_p * len(keep))(*[p.ctypes.data for p in keep])
    return keep, arr


def get_codes(planes, kinds, filt, idx):
    keep, arr = _plane_ptrs(planes)
    kinds = np.array([KIND[k] if isinstance(k, str) else k for k in kinds], np.int32)
    filt = np.ascontiguousarray(filt, np.uint8)
    idx = np.ascontiguousarray(idx, np.int64)
    codes = np.empty((idx.size, len(keep)), np.uint8)
    pf = np.empty(idx.size, np.uint8)
    lib().orc_get_codes(arr, _p(kinds), len(keep), _p(filt), C.c_uint32(filt.size), _p(idx), C.c_uint32(idx.size),
                        _p(codes), _p(pf))
    return codes, pf


def levenshtein(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_levenshtein(_p(a), a.size, _p(b), b.size)


def prefix_band_min(a, b, p, k):
    """Definition of the fused kernel's prefix test (not a reference function; welldup_oracle.c)."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_prefix_band_min(_p(a), a.size, _p(b), int(p), int(k))


def count_tile(planes, kinds, filt, centres, level_offsets, idx, levels, edit_distance=2, use_hamming=False,
               want_per_target=True):
    """planes: one array per compared position (repeat a plane to repeat a cycle)."""
    keep, arr = _plane_ptrs(planes)
    kinds = np.array([KIND[k] if isinstance(k, str) else k for k in kinds], np.int32)
    filt = np.ascontiguousarray(filt, np.uint8)
    centres = np.ascontiguousarray(centres, np.uint32)
    level_offsets = np.ascontiguousarray(level_offsets, np.uint32)
    idx = np.ascontiguousarray(idx, np.uint32)
    t = centres.size
    pt = np.zeros((t, 1 + 2 * levels), np.int32) if want_per_target else None
    counters = np.zeros(1 + 5 * levels, np.int64)
    lib().orc_count_tile(arr, _p(kinds), len(keep), _p(filt), C.c_uint32(filt.size), _p(centres), _p(level_offsets),
                         _p(idx), C.c_uint32(t), levels, edit_distance, 1 if use_hamming else 0,
                         _p(pt) if want_per_target else None, _p(counters))
    return pt, counters


def count_exhaustive(X, Y, planes, kinds, filt, levels=5, edit_distance=2, use_hamming=False):
    """Exhaustive mode: every well of the tile is a target (what
    prepare_cluster_indexes.py -n <wells> followed by count_well_duplicates.py
    computes; the counters do not depend on the sample order).  Raises
    RuntimeError when some well has an empty ring, as the reference would."""
    centres = np.arange(len(X), dtype=np.uint32)
    offs, idx = rings_csr(X, Y, centres, levels)
    _, counters = count_tile(planes, kinds, filt, centres, offs, idx, levels, edit_distance, use_hamming,
                             want_per_target=False)
    return counters
