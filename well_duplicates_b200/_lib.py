"""ctypes binding of libwelldup.so (include/welldup.h).

There is no CPU fallback: if the shared library has not been built, or no CUDA
device is usable, the functions here raise."""
import ctypes as C
import gzip
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwelldup.so")

WD_OK = 0
WD_E_INDEX, WD_E_RUNTIME, WD_E_ASSERT, WD_E_CUDA, WD_E_ARG, WD_E_CAPACITY = -1, -2, -3, -4, -5, -6
WD_E_NOENT, WD_E_IO, WD_E_EOF, WD_E_DATA = -7, -8, -9, -10
PLANE_EMPTY, PLANE_BCL, PLANE_CBCL, PLANE_CBCL_EXCL = 0, 1, 2, 3
MAX_LEVELS = 15
MAX_SEQ_LEN = 1024
ABI_VERSION = 2
MODE_FUSED, MODE_TWO_PASS, MODE_FUSED_LOG = 0, 1, 2
COMM_ID_BYTES = 128


class Tuning(C.Structure):
    """struct wd_tuning (include/welldup.h)."""
    _fields_ = [("step0", C.c_int32), ("step1", C.c_int32), ("centre_chunk", C.c_int32), ("head_planes", C.c_int32),
                ("head_groups", C.c_int32), ("visit_order", C.c_int32), ("targets_per_cta", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("reserved", C.c_int32 * 8)]


class CudaError(RuntimeError):
    """The CUDA runtime reported a failure (WD_E_CUDA)."""


class CapacityError(RuntimeError):
    """A caller-provided output array was too small (WD_E_CAPACITY)."""


# the staging errors are the ones gzip.open(...).read() raises in the reference's reader
_EXC = {WD_E_INDEX: IndexError, WD_E_RUNTIME: RuntimeError, WD_E_ASSERT: AssertionError,
        WD_E_CUDA: CudaError, WD_E_ARG: ValueError, WD_E_CAPACITY: CapacityError,
        WD_E_NOENT: FileNotFoundError, WD_E_IO: OSError, WD_E_EOF: EOFError, WD_E_DATA: gzip.BadGzipFile}


class InflateJob(C.Structure):
    """struct wd_inflate_job (include/welldup.h)."""
    _fields_ = [("path", C.c_char_p), ("src", C.c_void_p), ("offset", C.c_uint64), ("size", C.c_uint64),
                ("dst", C.c_void_p), ("dst_cap", C.c_uint64), ("out_len", C.c_uint64), ("status", C.c_int32),
                ("message", C.c_char * 220)]


_p = C.c_void_p
_u8p, _i32p, _u32p, _i64p, _u64p, _f32p = (C.POINTER(t) for t in
                                           (C.c_uint8, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_float))
SIGNATURES = {
    "wd_abi_version": (C.c_int, []),
    "wd_last_error": (C.c_char_p, []),
    "wd_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "wd_destroy": (C.c_int, [_p]),
    "wd_set_stream": (C.c_int, [_p, _p]),
    "wd_sync": (C.c_int, [_p]),
    "wd_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_p)]),
    "wd_host_free": (C.c_int, [_p]),
    "wd_host_register": (C.c_int, [_p, C.c_size_t]),
    "wd_host_unregister": (C.c_int, [_p]),
    "wd_launch_count": (C.c_int, [_p, _u64p]),
    "wd_last_count_h2d_bytes": (C.c_int, [_p, _u64p]),
    "wd_last_count_staging": (C.c_int, [_p, _i32p, C.POINTER(C.c_double)]),
    "wd_set_l2_fetch_granularity": (C.c_int, [_p, C.c_int, _i32p]),
    "wd_locs_load": (C.c_int, [_p, _p, C.c_uint32]),
    "wd_locs_pixels": (C.c_int, [_p, _p, _p]),
    "wd_ring_query": (C.c_int, [_p, _p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, _p, _p, C.c_size_t,
                                _u64p, _u32p]),
    "wd_targets_load": (C.c_int, [_p, _p, _p, _p, C.c_uint32, C.c_int]),
    "wd_tile_begin": (C.c_int, [_p, C.c_int, C.c_uint32, C.c_int]),
    "wd_tile_put_filter": (C.c_int, [_p, C.c_int, _p, C.c_uint32]),
    "wd_tile_put_bcl": (C.c_int, [_p, C.c_int, C.c_int, _p, C.c_uint32]),
    "wd_tile_put_cbcl": (C.c_int, [_p, C.c_int, C.c_int, _p, C.c_uint32, C.c_uint32, C.c_int]),
    "wd_tile_map_host": (C.c_int, [_p, C.c_int, C.c_uint32, C.c_int, _p, C.c_size_t, _p, _p, _p]),
    "wd_filter_offsets": (C.c_int, [_p, C.c_int, _p, _u32p]),
    "wd_get_seqs": (C.c_int, [_p, C.c_int, _p, C.c_uint32, _p, C.c_int, _p, _p]),
    "wd_count": (C.c_int, [_p, C.c_int, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "wd_count_async": (C.c_int, [_p, C.c_int, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "wd_count_fetch": (C.c_int, [_p, _p, _p]),
    "wd_dup_pairs": (C.c_int, [_p, _p, C.c_size_t, _u64p]),
    "wd_dup_pairs_seqs": (C.c_int, [_p, _p, _p, C.c_size_t, _u64p]),
    "wd_count_trace_sectors": (C.c_int, [_p, C.c_int, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "wd_set_tuning": (C.c_int, [_p, _p]),
    "wd_comm_unique_id": (C.c_int, [_p]),
    "wd_comm_init": (C.c_int, [_p, _p, C.c_int, C.c_int]),
    "wd_comm_destroy": (C.c_int, [_p]),
    "wd_allreduce_i64": (C.c_int, [_p, _p, C.c_size_t]),
    "wd_published_fetch": (C.c_int, [_p, _p, C.c_size_t]),
    "wd_comm_join": (C.c_int, [_p]),
    "wd_publish_add": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.POINTER(_p), C.POINTER(C.c_size_t)]),
    "wd_publish_counters": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.POINTER(_p), C.POINTER(C.c_size_t)]),
    "wd_counters_devptr": (C.c_int, [_p, C.POINTER(_p), C.POINTER(C.c_size_t)]),
    "wd_inflate_batch": (C.c_int, [C.POINTER(InflateJob), C.c_size_t, C.c_int]),
    "wd_gunzip": (C.c_int, [_p, C.c_size_t, _p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "wd_crc32": (C.c_uint32, [C.c_uint32, _p, C.c_size_t]),
    "wd_count_exhaustive": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _p]),
}

_lib = None


def load():
    """dlopen the library (once).  Raises if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m well_duplicates_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.wd_abi_version() != ABI_VERSION:
        raise ImportError("libwelldup.so has ABI version %d, expected %d" % (lib.wd_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def raise_status(rc, msg):
    raise _EXC.get(rc, RuntimeError)(msg)


def check(rc):
    if rc == WD_OK:
        return
    raise_status(rc, load().wd_last_error().decode("utf-8", "replace"))
