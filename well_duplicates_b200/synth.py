"""Synthetic patterned-flowcell generator.

Writes the on-disk formats the hot path consumes, to the layouts the reference
reads them from (no reference code is used here; the citations say where the
reference *parses* each format):

* ``s.locs``       12-byte header ``<ifI`` + N x (f32 x, f32 y)
                   (prepare_cluster_indexes.py:135-137, :99-116)
* ``*.filter``     ``<III`` = (0, 3, N) + N bytes, bit0 = pass-filter
                   (bcl_direct_reader.py:146-152, :222-253)
* ``*.bcl.gz``     gzip of ``<I`` N + N bytes, bits0-1 base, bits2-7 quality,
                   byte 0 = no-call (bcl_direct_reader.py:327-361)
* ``*.cbcl``       header/tile table/gzip members of 4-bit calls, with the
                   "non-PF clusters excluded" flag (cbcl_read.py:20-84,
                   bcl_direct_reader.py:255-325)

Lattice constants follow plan.md:28-66 of the reference (HiSeq 4000: row
length 1571, x pitch ~20.35 px, y pitch ~17.62 px, odd rows shifted by half a
pitch).  Everything is driven by a ``numpy.random.Generator`` so a seed fixes
the flowcell.
"""
from __future__ import annotations

import gzip
import os
import struct
from dataclasses import dataclass

import numpy as np

HISEQ4000_ROW_LEN = 1571
HISEQ4000_WELLS = 4309650          # bcl_direct_reader.py:144-145
NOVASEQ_WELLS = 4091904            # cbcl_read.py:77-78
X_PITCH = 20.35
Y_PITCH = 17.62
QUALS = np.array([7, 12, 23, 27, 32, 37, 41], dtype=np.uint8)


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def hex_lattice(n_wells: int, row_len: int, x_pitch: float = X_PITCH,
                y_pitch: float = Y_PITCH, x0: int = 1200, y0: int = 1100):
    """Integer pixel coordinates (X, Y) of a raster-ordered hex lattice.

    Returns int32 arrays of length ``n_wells``; the last row may be partial.
    """
    idx = np.arange(n_wells, dtype=np.int64)
    row = idx // row_len
    col = idx % row_len
    x = x0 + np.floor(col * x_pitch + (row & 1) * (x_pitch / 2.0) + 0.5)
    y = y0 + np.floor(row * y_pitch + 0.5)
    return x.astype(np.int32), y.astype(np.int32)


def xy_to_locs_floats(X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """f32 pairs that decode back to (X, Y) through int(f*10.0 + 1000.5)."""
    xy = np.empty((X.size, 2), dtype=np.float32)
    xy[:, 0] = (X.astype(np.float64) - 1000.0) / 10.0
    xy[:, 1] = (Y.astype(np.float64) - 1000.0) / 10.0
    back = (xy.astype(np.float64) * 10.0 + 1000.5).astype(np.int64)
    if not (np.array_equal(back[:, 0], X) and np.array_equal(back[:, 1], Y)):
        raise ValueError("coordinates do not survive the f32 round trip")
    return xy


def locs_bytes(xy: np.ndarray) -> bytes:
    xy = np.ascontiguousarray(xy, dtype="<f4")
    return struct.pack("<ifI", 1, 1.0, xy.shape[0]) + xy.tobytes()


def write_locs(path: str, xy: np.ndarray) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(locs_bytes(xy))


# --------------------------------------------------------------------------
# base calls
# --------------------------------------------------------------------------
@dataclass
class TileData:
    """One tile's raw inputs as the host sees them after gunzip."""
    planes: np.ndarray          # uint8 [cycles, N]  BCL bytes
    filt: np.ndarray            # uint8 [N]          filter bytes (bit0 = PF)

    @property
    def n_wells(self) -> int:
        return int(self.filt.size)

    @property
    def n_cycles(self) -> int:
        return int(self.planes.shape[0])


def lattice_offsets(row_len: int, rings: int = 5):
    """Index offsets of wells within ``rings`` rows/cols of a centre."""
    offs = []
    for dr in range(-rings, rings + 1):
        for dc in range(-rings, rings + 1):
            if dr or dc:
                offs.append(dr * row_len + dc)
    return np.array(offs, dtype=np.int64)


def make_tile(rng: np.random.Generator, n_wells: int, n_cycles: int, row_len: int,
              pf_rate: float = 0.72, nocall_rate: float = 0.005,
              dup_rate: float = 0.01, shift_share: float = 0.25,
              max_subs: int = 3) -> TileData:
    """Uniform random calls with planted local duplicates.

    ``dup_rate`` of the wells copy the read of a well up to 5 lattice steps
    away, with 0..max_subs substitutions; ``shift_share`` of those copies are
    shifted by one base instead (Levenshtein 2, Hamming large), so the default
    (Levenshtein) and ``--hamming`` modes give different counts.
    """
    base = np.empty((n_cycles, n_wells), dtype=np.uint8)
    for c in range(n_cycles):
        base[c] = rng.integers(0, 4, size=n_wells, dtype=np.uint8)
    n_dup = int(n_wells * dup_rate)
    if n_dup:
        offs = lattice_offsets(row_len)
        dst = rng.choice(n_wells, size=n_dup, replace=False)
        src = dst + offs[rng.integers(0, offs.size, size=n_dup)]
        ok = (src >= 0) & (src < n_wells)
        dst, src = dst[ok], src[ok]
        shifted = rng.random(dst.size) < shift_share
        # plain copies (+ substitutions)
        d0, s0 = dst[~shifted], src[~shifted]
        base[:, d0] = base[:, s0]
        nsub = rng.integers(0, max_subs + 1, size=d0.size)
        for k in range(1, max_subs + 1):
            sel = d0[nsub >= k]
            cyc = rng.integers(0, n_cycles, size=sel.size)
            base[cyc, sel] = (base[cyc, sel] + rng.integers(1, 4, size=sel.size, dtype=np.uint8)) & 3
        # one-base shifts
        d1, s1 = dst[shifted], src[shifted]
        if n_cycles > 1 and d1.size:
            base[1:, d1] = base[:-1, s1]
    planes = base            # finished in place, one cycle at a time (bounded memory at full tile size)
    for c in range(n_cycles):
        qual = QUALS[rng.integers(0, QUALS.size, size=n_wells, dtype=np.uint8)]
        planes[c] |= qual << 2
        if nocall_rate > 0:
            planes[c][rng.random(n_wells, dtype=np.float32) < nocall_rate] = 0
    filt = (rng.random(n_wells) < pf_rate).astype(np.uint8)
    # real filter files sometimes carry other bits; bit0 alone decides PF
    filt |= (rng.integers(0, 2, size=n_wells, dtype=np.uint8) << 1)
    return TileData(planes=planes, filt=filt)


def filter_file_bytes(filt: np.ndarray) -> bytes:
    return struct.pack("<III", 0, 3, filt.size) + np.ascontiguousarray(filt, np.uint8).tobytes()


def bcl_plane_bytes(plane: np.ndarray) -> bytes:
    return struct.pack("<I", plane.size) + np.ascontiguousarray(plane, np.uint8).tobytes()


def bcl_to_nibbles(plane: np.ndarray) -> np.ndarray:
    """BCL bytes -> CBCL 4-bit codes (2-bit quality bin | 2-bit base); 0 = no-call."""
    q = plane >> 2
    qbin = (1 + (q > 20).astype(np.uint8) + (q > 30).astype(np.uint8)).astype(np.uint8)
    nib = ((qbin << 2) | (plane & 3)).astype(np.uint8)
    nib[plane == 0] = 0
    return nib


def pack_nibbles(nib: np.ndarray) -> np.ndarray:
    """Two wells per byte, even well in the low half (cbcl_read.py:149-180)."""
    if nib.size & 1:
        nib = np.concatenate([nib, np.zeros(1, np.uint8)])
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


def basecalls_dir(run_root: str, lane: int) -> str:
    return os.path.join(run_root, "Data", "Intensities", "BaseCalls", "L%03d" % lane)


def write_bcl_tile(run_root: str, lane: int, tile: int, data: TileData,
                   first_cycle: int = 0, compresslevel: int = 1) -> None:
    """``L00x/s_<lane>_<tile>.filter`` + ``L00x/C<n>.1/s_<lane>_<tile>.bcl.gz``."""
    ldir = basecalls_dir(run_root, lane)
    os.makedirs(ldir, exist_ok=True)
    stem = "s_%d_%d" % (lane, tile)
    with open(os.path.join(ldir, stem + ".filter"), "wb") as fh:
        fh.write(filter_file_bytes(data.filt))
    for c in range(data.n_cycles):
        cdir = os.path.join(ldir, "C%d.1" % (first_cycle + c + 1))
        os.makedirs(cdir, exist_ok=True)
        with gzip.open(os.path.join(cdir, stem + ".bcl.gz"), "wb", compresslevel=compresslevel) as fh:
            fh.write(bcl_plane_bytes(data.planes[c]))


def cbcl_file_bytes(tiles: "dict[int, np.ndarray]", filters: "dict[int, np.ndarray]",
                    excluded: bool, compresslevel: int = 1) -> bytes:
    """One CBCL file (one lane-surface, one cycle).

    ``tiles[tile_no]`` is that tile's BCL-byte plane for this cycle (all wells);
    with ``excluded`` only PF wells are stored and the tile record's cluster
    count is the PF count (cbcl_read.py:77-80, :130-131).
    """
    members = []
    records = []
    for tno in sorted(tiles):
        nib = bcl_to_nibbles(tiles[tno])
        if excluded:
            nib = nib[(filters[tno] & 1).astype(bool)]
        packed = pack_nibbles(nib).tobytes()
        comp = gzip.compress(packed, compresslevel=compresslevel)
        members.append(comp)
        records.append(struct.pack("<IIII", tno, nib.size, len(packed), len(comp)))
    n_bins = 4
    header_size = 12 + n_bins * 8 + 4 + 16 * len(records) + 1
    out = [struct.pack("<HIBBI", 1, header_size, 2, 2, n_bins)]
    for frm, to in ((0, 0), (1, 11), (2, 25), (3, 37)):
        out.append(struct.pack("<II", frm, to))
    out.append(struct.pack("<I", len(records)))
    out.extend(records)
    out.append(bytes([1 if excluded else 0]))
    out.extend(members)
    return b"".join(out)


def write_cbcl_lane(run_root: str, lane: int, tiles: "dict[int, TileData]",
                    excluded_from_cycle: int, compresslevel: int = 1) -> None:
    """NovaSeq-style lane: ``L00x/s_<lane>_<tile>.filter`` and
    ``L00x/C<n>.1/L00x_<surface>.cbcl``; cycles >= ``excluded_from_cycle``
    (0-based) are written with the excluded flag set."""
    ldir = basecalls_dir(run_root, lane)
    os.makedirs(ldir, exist_ok=True)
    n_cycles = None
    for tno, td in tiles.items():
        with open(os.path.join(ldir, "s_%d_%d.filter" % (lane, tno)), "wb") as fh:
            fh.write(filter_file_bytes(td.filt))
        n_cycles = td.n_cycles if n_cycles is None else n_cycles
        assert td.n_cycles == n_cycles
    surfaces = sorted({str(t)[0] for t in tiles})
    for c in range(n_cycles):
        cdir = os.path.join(ldir, "C%d.1" % (c + 1))
        os.makedirs(cdir, exist_ok=True)
        for s in surfaces:
            sel = {t: td.planes[c] for t, td in tiles.items() if str(t)[0] == s}
            flt = {t: td.filt for t, td in tiles.items() if str(t)[0] == s}
            blob = cbcl_file_bytes(sel, flt, excluded=(c >= excluded_from_cycle),
                                   compresslevel=compresslevel)
            with open(os.path.join(cdir, "L%03d_%s.cbcl" % (lane, s)), "wb") as fh:
                fh.write(blob)


# --------------------------------------------------------------------------
# fast generator for benchmark-sized flowcells
# --------------------------------------------------------------------------
def _bcl_lut():
    """random byte -> BCL byte: bits0-1 base, bits2-4 pick a quality, 1 value in
    256 (0.4 %) is a no-call."""
    r = np.arange(256, dtype=np.uint16)
    q = np.concatenate([QUALS, QUALS[-1:]])[(r >> 2) & 7]
    lut = ((r & 3) | (q.astype(np.uint16) << 2)).astype(np.uint8)
    lut[0xA5] = 0
    return lut


def make_tile_fast(seed: int, n_wells: int, n_cycles: int, row_len: int, pf_rate: float = 0.72,
                   dup_rate: float = 0.01, shift_share: float = 0.25, out: "np.ndarray | None" = None) -> TileData:
    """Same shape of data as make_tile (uniform calls, ~0.4 % no-calls, planted
    substitution and one-base-shift duplicates) at a few hundred MB/s, written
    straight into ``out`` ([cycles, N] uint8, e.g. pinned memory) when given."""
    rng = np.random.default_rng(seed)
    planes = out if out is not None else np.empty((n_cycles, n_wells), dtype=np.uint8)
    lut = _bcl_lut()
    for c in range(n_cycles):
        raw = np.frombuffer(rng.bytes(n_wells), dtype=np.uint8)
        np.take(lut, raw, out=planes[c])
    n_dup = int(n_wells * dup_rate)
    if n_dup:
        offs = lattice_offsets(row_len)
        dst = rng.choice(n_wells, size=n_dup, replace=False)
        src = dst + offs[rng.integers(0, offs.size, size=n_dup)]
        ok = (src >= 0) & (src < n_wells)
        dst, src = dst[ok], src[ok]
        shifted = rng.random(dst.size) < shift_share
        d0, s0 = dst[~shifted], src[~shifted]
        planes[:, d0] = planes[:, s0]
        nsub = rng.integers(0, 4, size=d0.size)
        for k in range(1, 4):
            sel = d0[nsub >= k]
            cyc = rng.integers(0, n_cycles, size=sel.size)
            planes[cyc, sel] ^= rng.integers(1, 4, size=sel.size, dtype=np.uint8)   # changes the base bits only
        d1, s1 = dst[shifted], src[shifted]
        if n_cycles > 1 and d1.size:
            planes[1:, d1] = planes[:-1, s1]
    filt = (rng.random(n_wells) < pf_rate).astype(np.uint8)
    return TileData(planes=planes, filt=filt)
