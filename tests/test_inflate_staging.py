"""Host staging pipeline: the library's gunzip (wd_gunzip /
wd_inflate_batch / wd_crc32, csrc/wd_inflate.cc) against zlib / gzip -- what
the reference's reader calls (bcl_direct_reader.py:207-208, :300-301) -- and
staging.Stager against a plain-Python reading of the same run folders.
Host-only: runs without a GPU (pageable blocks instead of page-locked ones)."""
import ctypes as C
import gzip
import io
import os
import random
import struct
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from well_duplicates_b200 import _lib, reader, staging, synth


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def native_gunzip(lib, data, cap):
    out = np.full(cap + 64, 0xAA, np.uint8)
    n = C.c_size_t()
    rc = lib.wd_gunzip(data, len(data), out.ctypes.data, cap, C.byref(n))
    assert (out[cap:] == 0xAA).all(), "wrote past the capacity it was given"
    return rc, out[:n.value].tobytes()


def deflate(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return c.compress(raw) + c.flush()


def bcl_like(rng, n, quals=(7, 12, 23, 27, 32, 37, 41), nocall=0.005):
    q = np.array(quals, np.uint8) * 4
    b = q[rng.integers(0, len(quals), n)] | rng.integers(0, 4, n).astype(np.uint8)
    b[rng.random(n) < nocall] = 0
    return b.tobytes()


def payloads():
    rnd = random.Random(1)
    rng = np.random.default_rng(3)
    yield "empty", b""
    yield "one byte", b"a"
    yield "short text", b"hello world" * 3
    yield "zeros (distance 1, longest matches)", bytes(70000)
    yield "random bytes (stored blocks)", rnd.randbytes(70000)
    yield "base calls", bcl_like(rng, 300000)
    yield "two-bit symbols", bytes(rnd.getrandbits(8) & 3 for _ in range(100000))
    yield "period 3", b"abc" * 30000
    yield "short periods", b"".join(bytes([i % 251]) * (i % 300) for i in range(1200))
    yield "all byte values", bytes(range(256)) * 300
    # skewed alphabets give code lengths beyond the first-level table
    yield "geometric", np.random.default_rng(5).geometric(0.08, 300000).clip(0, 255).astype(np.uint8).tobytes()
    yield "zipf", (np.random.default_rng(6).zipf(1.3, 300000) % 256).astype(np.uint8).tobytes()
    nib = bcl_like(rng, 200000, quals=(2, 12, 23, 37))
    yield "cbcl nibbles", synth.pack_nibbles(synth.bcl_to_nibbles(np.frombuffer(nib, np.uint8))).tobytes()


def test_crc32_equals_zlib(lib):
    rnd = random.Random(2)
    for n in list(range(0, 300)) + [1000, 4096, 65537, (1 << 20) + 5]:
        b = rnd.randbytes(n)
        for init in (0, 0x12345678, 0xFFFFFFFF):
            assert lib.wd_crc32(init, b, n) == zlib.crc32(b, init), (n, init)
    # running value over pieces
    b = rnd.randbytes(100000)
    c = 0
    for k in range(0, len(b), 777):
        c = lib.wd_crc32(c, b[k:k + 777], len(b[k:k + 777]))
    assert c == zlib.crc32(b)


@pytest.mark.parametrize("name,raw", list(payloads()), ids=[n for n, _ in payloads()])
def test_gunzip_equals_zlib_for_every_block_type(lib, name, raw):
    for level in (0, 1, 6, 9):                  # 0: stored blocks
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED, zlib.Z_FILTERED):
            comp = deflate(raw, level, strategy)
            rc, out = native_gunzip(lib, comp, len(raw))
            assert rc == 0 and out == raw, (level, strategy, lib.wd_last_error())
            if raw:
                rc, out = native_gunzip(lib, comp, len(raw) - 1)          # one byte too small
                assert rc == _lib.WD_E_CAPACITY and out == raw[:-1]
            for cut in (1, 5, 9, 12, len(comp) // 2, len(comp) - 9, len(comp) - 8, len(comp) - 1):
                if 0 < cut < len(comp):                                    # truncated file: EOFError in gzip.py
                    rc, out = native_gunzip(lib, comp[:cut], len(raw))
                    assert rc == _lib.WD_E_EOF, (cut, len(comp), level, strategy)
                    assert raw.startswith(out)


def test_gzip_framing_like_gzip_open(lib):
    rnd = random.Random(4)
    a, b = rnd.randbytes(5000), b"xyz" * 1000
    two = deflate(a) + deflate(b)
    assert native_gunzip(lib, two, 8000) == (0, a + b)                       # members follow each other
    assert native_gunzip(lib, two + bytes(700), 8000) == (0, a + b)          # zero padding after the last member
    assert native_gunzip(lib, deflate(a) + bytes(13) + deflate(b) + bytes(2), 8000) == (0, a + b)
    assert native_gunzip(lib, two + b"junk", 8000)[0] == _lib.WD_E_DATA     # gzip.BadGzipFile
    assert native_gunzip(lib, b"", 10) == (0, b"")                           # empty file reads as b""
    assert native_gunzip(lib, b"\0\0" + two, 8000)[0] == _lib.WD_E_DATA
    assert native_gunzip(lib, b"\x1f", 10)[0] == _lib.WD_E_EOF
    assert native_gunzip(lib, b"\x1f\x8b\x07" + bytes(20), 10)[0] == _lib.WD_E_DATA   # unknown method
    # a file written by gzip.GzipFile carries a name
    bio = io.BytesIO()
    with gzip.GzipFile(filename="s_1_1101.bcl", mode="wb", fileobj=bio, mtime=5) as fh:
        fh.write(a)
    assert native_gunzip(lib, bio.getvalue(), len(a)) == (0, a)
    # every optional header field at once
    body = deflate(a, wbits=-15)
    hdr = (b"\x1f\x8b\x08" + bytes([4 | 8 | 16 | 2]) + bytes(4) + b"\0\3" + struct.pack("<H", 5) + b"EXTRA" +
           b"name\0" + b"comment\0" + b"\x12\x34")
    full = hdr + body + struct.pack("<II", zlib.crc32(a), len(a))
    assert gzip.decompress(full) == a
    assert native_gunzip(lib, full, len(a)) == (0, a)
    for cut in range(4, len(hdr)):
        assert native_gunzip(lib, full[:cut], len(a))[0] == _lib.WD_E_EOF
    bad = bytearray(full)
    bad[-5] ^= 1
    assert native_gunzip(lib, bytes(bad), len(a))[0] == _lib.WD_E_DATA and b"CRC" in lib.wd_last_error()
    bad = bytearray(full)
    bad[-1] ^= 1
    assert native_gunzip(lib, bytes(bad), len(a))[0] == _lib.WD_E_DATA and b"length" in lib.wd_last_error()


def test_corrupt_streams_fail_like_zlib_and_stay_in_bounds(lib):
    rnd = random.Random(7)
    raw = bytes(rnd.choice(b"ACGTN") for _ in range(30000)) + rnd.randbytes(3000) + bytes(5000)
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (9, zlib.Z_FILTERED)):
        base = deflate(raw, level, strategy)
        for _ in range(700):
            bb = bytearray(base)
            for _ in range(rnd.randint(1, 4)):
                bb[rnd.randrange(10, len(bb))] = rnd.getrandbits(8)
            rc, out = native_gunzip(lib, bytes(bb), 40000)
            try:
                want = gzip.decompress(bytes(bb))
            except Exception:
                assert rc != 0
            else:
                assert rc == 0 and out == want
    # arbitrary bits as a deflate stream: any verdict, no crash, no write outside the buffer
    for _ in range(2000):
        native_gunzip(lib, b"\x1f\x8b\x08\0" + bytes(6) + rnd.randbytes(rnd.randint(1, 200)), 5000)


def test_gunzip_maps_to_the_exceptions_of_gzip_open():
    raw = b"ACGT" * 1000
    comp = deflate(raw)
    assert reader.gunzip(comp) == raw
    assert reader.gunzip(comp, size_hint=len(raw)) == raw
    assert reader.gunzip(deflate(bytes(10 << 20))) == bytes(10 << 20)          # grows past the first guess
    with pytest.raises(EOFError):
        reader.gunzip(comp[:-3])
    with pytest.raises(gzip.BadGzipFile):
        reader.gunzip(b"not a gzip file")
    with pytest.raises(OSError):                                              # BadGzipFile is an OSError, as in gzip.py
        reader.gunzip(comp[:20] + b"\xff\xff\xff\xff" + comp[24:])


def test_inflate_batch_files_offsets_and_statuses(lib, tmp_path):
    rng = np.random.default_rng(11)
    raws = [bcl_like(rng, 50000 + 1000 * k) for k in range(12)]
    # files 0..5 hold one member; file "multi" holds members 6..11 back to back at known offsets (a CBCL body)
    offs, blob = [], b"HEADER--"
    for r in raws[6:]:
        offs.append(len(blob))
        blob += deflate(r, 1)
    (tmp_path / "multi.cbcl").write_bytes(blob)
    for k in range(6):
        (tmp_path / ("f%d.bcl.gz" % k)).write_bytes(deflate(raws[k], 6))
    jobs = (_lib.InflateJob * 16)()
    outs = []
    keep = []
    for k in range(12):
        out = np.full(len(raws[k]) + 32, 0x55, np.uint8)
        outs.append(out)
        if k < 6:
            path, off, size = tmp_path / ("f%d.bcl.gz" % k), 0, 0
        else:
            path, off = tmp_path / "multi.cbcl", offs[k - 6]
            size = (offs[k - 5] if k < 11 else len(blob)) - off
        keep.append(os.fsencode(str(path)))
        jobs[k].path, jobs[k].offset, jobs[k].size = keep[-1], off, size
        jobs[k].dst, jobs[k].dst_cap = out.ctypes.data, len(raws[k])
    # failures: missing file, truncated file, not gzip, too small a destination
    (tmp_path / "cut.gz").write_bytes(deflate(raws[0])[:-20])
    (tmp_path / "text.gz").write_bytes(b"plain text, not gzip")
    small = np.zeros(100, np.uint8)
    for k, (name, cap) in enumerate((("missing.gz", 10), ("cut.gz", 60000), ("text.gz", 100), ("f0.bcl.gz", 100)), start=12):
        keep.append(os.fsencode(str(tmp_path / name)))
        jobs[k].path, jobs[k].dst, jobs[k].dst_cap = keep[-1], small.ctypes.data, min(cap, 100) if name != "cut.gz" else 0
    big = np.zeros(60000, np.uint8)
    jobs[13].dst, jobs[13].dst_cap = big.ctypes.data, big.size
    for threads in (1, 3, 0):
        for k in range(12):
            outs[k][:] = 0x55
        rc = lib.wd_inflate_batch(jobs, 16, threads)
        assert rc == _lib.WD_E_NOENT                                  # status of the first failed job
        assert b"missing.gz" in lib.wd_last_error()
        for k in range(12):
            assert jobs[k].status == 0 and jobs[k].out_len == len(raws[k])
            assert outs[k][:len(raws[k])].tobytes() == raws[k] and (outs[k][len(raws[k]):] == 0x55).all()
        assert [jobs[k].status for k in range(12, 16)] == [_lib.WD_E_NOENT, _lib.WD_E_EOF, _lib.WD_E_DATA, _lib.WD_E_CAPACITY]
        assert jobs[15].out_len == 100 and small.tobytes() == raws[0][:100]
    assert lib.wd_inflate_batch(jobs, 12, 2) == 0
    assert lib.wd_inflate_batch(None, 0, 4) == 0
    # memory source instead of a file
    job = (_lib.InflateJob * 1)()
    comp = deflate(raws[3])
    buf = np.frombuffer(comp, np.uint8)
    out = np.zeros(len(raws[3]), np.uint8)
    job[0].src, job[0].size, job[0].dst, job[0].dst_cap = buf.ctypes.data, len(comp), out.ctypes.data, out.size
    assert lib.wd_inflate_batch(job, 1, 1) == 0 and out.tobytes() == raws[3]


# ---- Stager against a plain reading of the same folders ---------------------------------------------
def plain_read(tile, cycle):
    """What bcl_direct_reader.py does with the files of one tile and cycle."""
    cdir = os.path.join(tile.data_dir, "C%d.1" % (cycle + 1))
    try:
        with gzip.open(os.path.join(cdir, tile.bcl_filename), "rb") as fh:
            raw = fh.read()
        assert struct.unpack("<I", raw[:4])[0] == tile.num_clusters
        return "bcl", np.frombuffer(raw, np.uint8, offset=4), tile.num_clusters
    except FileNotFoundError:
        cf = reader.CbclFile(os.path.join(cdir, tile.cbcl_filename))
        off, ncl, usize, csize = cf.blocks[int(tile.tile)]
        with open(cf.path, "rb") as fh:
            fh.seek(off)
            data = gzip.GzipFile(fileobj=fh, mode="rb").read(usize)
        return ("cbcl_excl" if cf.excluded else "cbcl"), np.frombuffer(data, np.uint8), ncl


def check_batch(batch, tiles, cycles):
    kind_name = {_lib.PLANE_BCL: "bcl", _lib.PLANE_CBCL: "cbcl", _lib.PLANE_CBCL_EXCL: "cbcl_excl"}
    assert batch.cycles == sorted(set(cycles))
    for k, t in enumerate(tiles):
        assert np.array_equal(batch.filter(k), t.read_filter())
        planes = batch.planes(k)
        for cyc in set(cycles):
            p = batch.plane_of[cyc]
            kind, data, ncl = plain_read(t, cyc)
            assert kind_name[batch.kinds[k, p]] == kind and batch.n_block[k, p] == ncl
            assert np.array_equal(planes[p, :data.size], data)


def test_stager_reads_golden_run_folders(manifest):
    st = staging.Stager(pinned=False, threads=3)
    for run, lane in (("run_bcl", 2), ("run_cbcl", 1)):
        rd = reader.BCLReader(os.path.join(GOLDEN, run))
        ldir = os.path.join(GOLDEN, run, "Data", "Intensities", "BaseCalls", "L%03d" % lane)
        names = sorted({m.group(1) for m in (__import__("re").match(r"s_\d_(\d+)\.filter", f) for f in os.listdir(ldir)) if m})
        assert names
        tiles = [rd.get_tile(lane, t) for t in names]
        ncyc = tiles[0].num_cycles
        groups = {}
        for t in tiles:
            groups.setdefault(t.num_clusters, []).append(t)
        for same in groups.values():
            for cycles in (list(range(ncyc)), [ncyc - 1, 0, 0, 3], []):
                for which in (0, 1):
                    check_batch(st.load(same, cycles, which), same, cycles)
    st.close()


def test_lane_batches_walk_a_lane_with_mixed_sizes_and_formats(tmp_path):
    rng = np.random.default_rng(8)
    run = str(tmp_path / "run")
    sizes = {1101: 3001, 1102: 3001, 1103: 2000, 1104: 2000, 1105: 2000, 1106: 777}
    for tile, n in sizes.items():
        synth.write_bcl_tile(run, 1, tile, synth.make_tile(rng, n, 6, 50), compresslevel=rng.integers(1, 9))
    cb = {t: synth.make_tile(rng, 1501, 6, 50) for t in (1101, 1102, 2101)}
    synth.write_cbcl_lane(run, 2, cb, excluded_from_cycle=3)
    rd = reader.BCLReader(run)
    st = staging.Stager(pinned=False, threads=2, cbcl_cache=rd._cbcl_cache)
    said = []
    for per_batch, want_groups in ((None, [[1101, 1102], [1103, 1104, 1105], [1106]]),
                                   (2, [[1101, 1102], [1103, 1104], [1105], [1106]]),
                                   (1, [[t] for t in sizes])):
        got = []
        for names, batch in staging.lane_batches(st, lambda t: rd.get_tile(1, t), list(sizes), [1, 2, 5, 2],
                                                 per_batch=per_batch, announce=said.append):
            got.append(names)
            check_batch(batch, [rd.get_tile(1, t) for t in names], [1, 2, 5, 2])
        assert got == want_groups
    assert said == list(sizes)                     # only the one-tile-at-a-time walk announces
    for names, batch in staging.lane_batches(st, lambda t: rd.get_tile(2, t), [1101, 2101, 1102], list(range(6))):
        assert names == [1101, 2101, 1102]
        check_batch(batch, [rd.get_tile(2, t) for t in names], list(range(6)))
        assert batch.kinds[0].tolist() == [_lib.PLANE_CBCL] * 3 + [_lib.PLANE_CBCL_EXCL] * 3
        assert batch.compressed_bytes > 0 and batch.inflated_bytes == int(batch.usize.sum())
    assert list(staging.lane_batches(st, lambda t: rd.get_tile(1, t), [], [0])) == []
    st.close()


def test_stager_raises_what_the_reference_reader_raises(tmp_path):
    rng = np.random.default_rng(9)
    run = str(tmp_path / "run")
    synth.write_bcl_tile(run, 1, 1101, synth.make_tile(rng, 1000, 4, 50))
    synth.write_bcl_tile(run, 1, 1102, synth.make_tile(rng, 1000, 4, 50))
    ldir = synth.basecalls_dir(run, 1)
    rd = reader.BCLReader(run)
    st = staging.Stager(pinned=False, threads=2)
    tiles = [rd.get_tile(1, 1101), rd.get_tile(1, 1102)]
    st.load(tiles, [0, 1, 2, 3])
    # neither a BCL nor a CBCL file for the cycle: FileNotFoundError (bcl_direct_reader.py:214)
    os.remove(os.path.join(ldir, "C3.1", "s_1_1102.bcl.gz"))
    with pytest.raises(FileNotFoundError):
        st.load(tiles, [0, 1, 2, 3])
    st.load(tiles, [0, 1, 3])
    # truncated file
    path = os.path.join(ldir, "C2.1", "s_1_1101.bcl.gz")
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:-11])
    with pytest.raises(EOFError):
        st.load(tiles, [1])
    # CRC mismatch
    open(path, "wb").write(blob[:-8] + b"\0\0\0\0" + blob[-4:])
    with pytest.raises(gzip.BadGzipFile):
        st.load(tiles, [1])
    # cluster count in the BCL header differs from the filter's (bcl_direct_reader.py:338)
    other = synth.make_tile(rng, 999, 1, 50)
    with gzip.open(path, "wb") as fh:
        fh.write(synth.bcl_plane_bytes(other.planes[0]))
    with pytest.raises(AssertionError):
        st.load(tiles, [1])
    longer = synth.make_tile(rng, 1001, 1, 50)
    with gzip.open(path, "wb") as fh:
        fh.write(synth.bcl_plane_bytes(longer.planes[0]))
    with pytest.raises(AssertionError):
        st.load(tiles, [1])
    st.close()


def test_inflate_under_address_and_ub_sanitizers(tmp_path):
    """The decoder compiled with -fsanitize=address,undefined survives truncated and bit-flipped
    streams of every block type with exact-size heap buffers (tests/inflate_fuzz.cc)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "inflate_fuzz")
    # WD_INFLATE_NO_MULTIVERSION: the plain build of the decoder (what a CPU without BMI2 runs); the library the
    # other tests load picks its BMI2 build on this machine
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-DWD_INFLATE_NO_MULTIVERSION", "-o", exe,
           os.path.join(root, "tests", "inflate_fuzz.cc"), os.path.join(root, "well_duplicates_b200", "csrc", "wd_inflate.cc"),
           "-lpthread"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "asan" in built.stderr.lower():
        pytest.skip("sanitizer runtime not installed")
    assert built.returncode == 0, built.stderr
    rnd = random.Random(5)
    raw = bytes(rnd.choice(b"ACGTN") for _ in range(120000)) + rnd.randbytes(20000) + bytes(30000) + b"abc" * 9000
    files = []
    for name, level, strategy in (("l6", 6, zlib.Z_DEFAULT_STRATEGY), ("fixed", 6, zlib.Z_FIXED), ("stored", 0, zlib.Z_DEFAULT_STRATEGY),
                                  ("rle", 6, zlib.Z_RLE), ("huffman", 6, zlib.Z_HUFFMAN_ONLY)):
        files.append(str(tmp_path / (name + ".gz")))
        with open(files[-1], "wb") as fh:
            fh.write(deflate(raw, level, strategy))
    run = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "fuzz runs 2000" in run.stdout


# ---- deflate streams zlib's own compressor never writes -------------------------------------------
class _BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def bits(self, value, count):                 # LSB first (header fields, extra bits)
        self.acc |= value << self.n
        self.n += count
        while self.n >= 8:
            self.out.append(self.acc & 0xff)
            self.acc >>= 8
            self.n -= 8

    def code(self, code, length):                 # Huffman codes go MSB first
        self.bits(int(format(code, "0%db" % length)[::-1], 2), length)

    def done(self):
        if self.n:
            self.out.append(self.acc & 0xff)
        return bytes(self.out)


_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
              8193, 12289, 16385, 24577]
_DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def _huffman_lengths(freq, limit=15):
    """Code lengths of a Huffman code for the symbols with freq > 0 (complete by construction);
    frequencies are flattened until no code is longer than ``limit``."""
    import heapq
    freq = list(freq)
    while True:
        heap = [(f, i, (i,)) for i, f in enumerate(freq) if f > 0]
        lens = [0] * len(freq)
        if len(heap) == 1:
            lens[heap[0][1]] = 1
            return lens
        heapq.heapify(heap)
        tick = len(freq)
        while len(heap) > 1:
            a, b = heapq.heappop(heap), heapq.heappop(heap)
            for s in a[2] + b[2]:
                lens[s] += 1
            heapq.heappush(heap, (a[0] + b[0], tick, a[2] + b[2]))
            tick += 1
        if max(lens) <= limit:
            return lens
        freq = [(f + 1) // 2 + 1 if f > 0 else 0 for f in freq]


def _canonical(lens):
    codes, code = {}, 0
    for length in range(1, 16):
        for s, l in enumerate(lens):
            if l == length:
                codes[s] = (code, length)
                code += 1
        code <<= 1
    return codes


def _tokens(rnd, data, p_match):
    """Greedy-random LZ77 parse: any earlier occurrence within 32768 bytes, any length up to 258,
    including overlapping copies and distances of 1..15."""
    seen, pos, out = {}, 0, []
    while pos < len(data):
        key = data[pos:pos + 3]
        cands = [c for c in seen.get(key, ()) if pos - c <= 32768]
        if len(key) == 3 and cands and rnd.random() < p_match:
            src = rnd.choice(cands[-8:] + cands[:2])
            length = 3
            while length < 258 and pos + length < len(data) and data[src + length] == data[pos + length]:
                length += 1
            length = rnd.randint(3, length)
            out.append((length, pos - src))
        else:
            out.append(data[pos])
            length = 1
        for k in range(pos, pos + length):
            seen.setdefault(data[k:k + 3], []).append(k)
        pos += length
    return out


def _sym_of(value, base):
    s = 0
    while s + 1 < len(base) and base[s + 1] <= value:
        s += 1
    return s


def _encode_dynamic(w, tokens, final, skew):
    lit_freq, dist_freq = [0] * 286, [0] * 30
    for t in tokens:
        if isinstance(t, tuple):
            lit_freq[257 + _sym_of(t[0], _LEN_BASE)] += 1
            dist_freq[_sym_of(t[1], _DIST_BASE)] += 1
        else:
            lit_freq[t] += 1
    lit_freq[256] = 1
    if skew:                                       # stretch the code: rare symbols get the longest codes allowed
        lit_freq = [f ** 3 if f else 0 for f in lit_freq]
        dist_freq = [f ** 3 if f else 0 for f in dist_freq]
    if not any(dist_freq):
        dist_freq[0] = 1
    ll, dl = _huffman_lengths(lit_freq), _huffman_lengths(dist_freq)
    n_lit = max(257, max(i for i, l in enumerate(ll) if l) + 1)
    n_dist = max(1, max(i for i, l in enumerate(dl) if l) + 1)
    w.bits(1 if final else 0, 1)
    w.bits(2, 2)
    w.bits(n_lit - 257, 5)
    w.bits(n_dist - 1, 5)
    w.bits(19 - 4, 4)
    # code-length code: a complete code over the 19 symbols (13 of 4 bits, 6 of 5 bits), no run-length symbols used
    pre_lens = [5, 5, 5] + [4] * 13 + [5, 5, 5]    # symbols 16, 17, 18, 0..15 -> indexed by symbol below
    by_symbol = {16: 5, 17: 5, 18: 5}
    by_symbol.update({s: 4 for s in range(13)})
    by_symbol.update({13: 5, 14: 5, 15: 5})
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    for s in order:
        w.bits(by_symbol[s], 3)
    pre = _canonical([by_symbol[s] for s in range(19)])
    for l in ll[:n_lit] + dl[:n_dist]:
        w.code(*pre[l])
    lit, dist = _canonical(ll), _canonical(dl)
    for t in tokens:
        if isinstance(t, tuple):
            ls = _sym_of(t[0], _LEN_BASE)
            w.code(*lit[257 + ls])
            w.bits(t[0] - _LEN_BASE[ls], _LEN_EXTRA[ls])
            ds = _sym_of(t[1], _DIST_BASE)
            w.code(*dist[ds])
            w.bits(t[1] - _DIST_BASE[ds], _DIST_EXTRA[ds])
        else:
            w.code(*lit[t])
    w.code(*lit[256])
    del pre_lens


def _encode_fixed(w, tokens, final):
    ll = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
    lit, dist = _canonical(ll), {s: (s, 5) for s in range(30)}
    w.bits(1 if final else 0, 1)
    w.bits(1, 2)
    for t in tokens:
        if isinstance(t, tuple):
            ls = _sym_of(t[0], _LEN_BASE)
            w.code(*lit[257 + ls])
            w.bits(t[0] - _LEN_BASE[ls], _LEN_EXTRA[ls])
            ds = _sym_of(t[1], _DIST_BASE)
            w.code(*dist[ds])
            w.bits(t[1] - _DIST_BASE[ds], _DIST_EXTRA[ds])
        else:
            w.code(*lit[t])
    w.code(*lit[256])


def test_streams_from_a_foreign_encoder(lib):
    """Valid deflate that zlib's compressor never writes -- matches at the full 32768 distance and of
    every length, overlapping copies at distances 1..15, 15-bit codes, empty blocks of all three
    types in between, a match that crosses block boundaries' history -- decoded like zlib's inflate
    decodes it."""
    rnd = random.Random(12)
    alphabet = bytes(rnd.sample(range(256), 40))
    for trial in range(8):
        n = rnd.choice([3000, 40000, 90000])
        data = bytearray(rnd.choice(alphabet) for _ in range(n))
        for _ in range(n // 400):                   # long repeats, some of them 32768 back, some runs
            length = rnd.randint(3, 600)
            dst = rnd.randrange(0, max(1, n - length))
            back = rnd.choice([1, 2, 7, 15, 16, 17, 255, 4096, 32767, 32768])
            if dst - back >= 0:
                for k in range(length):
                    data[dst + k] = data[dst + k - back]
        data = bytes(data)
        tokens = _tokens(rnd, data, p_match=rnd.choice([0.3, 0.9, 1.0]))
        w = _BitWriter()
        cuts = sorted(rnd.sample(range(1, len(tokens)), min(4, len(tokens) - 1))) + [len(tokens)]
        start = 0
        for bi, end in enumerate(cuts):
            final = end == len(tokens)
            kind = rnd.choice(["dyn", "dyn_skew", "fixed"])
            if kind == "fixed":
                _encode_fixed(w, tokens[start:end], final and bi % 2 == 0)
            else:
                _encode_dynamic(w, tokens[start:end], final and bi % 2 == 0, kind == "dyn_skew")
            if not (final and bi % 2 == 0):
                # an empty stored block (what Z_SYNC_FLUSH writes), then maybe an empty fixed block
                w.bits(0, 1)
                w.bits(0, 2)
                if w.n:
                    w.bits(0, 8 - w.n)
                w.bits(0, 16)
                w.bits(0xffff, 16)
                if final:
                    w.bits(1, 1)
                    w.bits(1, 2)
                    w.code(0, 7)                    # end-of-block in the fixed code
            start = end
        raw_deflate = w.done()
        assert zlib.decompress(raw_deflate, wbits=-15) == data           # the stream is valid
        framed = b"\x1f\x8b\x08\0" + bytes(6) + raw_deflate + struct.pack("<II", zlib.crc32(data), len(data))
        rc, out = native_gunzip(lib, framed, len(data))
        assert rc == 0 and out == data, (trial, rc, lib.wd_last_error())
        rc, out = native_gunzip(lib, framed, len(data) - 1)
        assert rc == _lib.WD_E_CAPACITY and out == data[:-1]


def test_inflate_batch_under_thread_sanitizer(tmp_path):
    """wd_inflate_batch's native thread pool under -fsanitize=thread (tests/inflate_batch_tsan.cc)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "inflate_tsan")
    built = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-DWD_INFLATE_NO_MULTIVERSION", "-o", exe,
                            os.path.join(root, "tests", "inflate_batch_tsan.cc"),
                            os.path.join(root, "well_duplicates_b200", "csrc", "wd_inflate.cc"), "-lpthread"],
                           capture_output=True, text=True)
    if built.returncode != 0 and "tsan" in built.stderr.lower():
        pytest.skip("sanitizer runtime not installed")
    assert built.returncode == 0, built.stderr
    rng = np.random.default_rng(2)
    files, total = [], 0
    for k in range(5):
        raw = bcl_like(rng, 150000 + 1000 * k)
        files.append(str(tmp_path / ("p%d.gz" % k)))
        with open(files[-1], "wb") as fh:
            fh.write(deflate(raw, 1 + k))
    for k in range(64):
        total += 150000 + 1000 * (k % 5)
    run = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    if run.returncode != 0 and "unexpected memory mapping" in run.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert run.returncode == 0 and "WARNING: ThreadSanitizer" not in run.stderr, run.stdout + run.stderr
    assert "batch rc 0 total %d" % total in run.stdout
