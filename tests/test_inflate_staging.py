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
