"""Builds libwelldup.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwelldup.so")
SOURCES = ["wd_inst23_w16.cu", "wd_inst23_w8.cu", "wd_inst23_w4.cu", "wd_inst23_w2.cu", "wd_inst23_w1.cu",
           "wd_api.cu", "wd_stage1.cu", "wd_stage23.cu", "wd_exhaustive.cu",
           "wd_inflate.cc",      # host-only: gunzip of the staging pipeline
           "wd_comm.cc",         # host-only: NCCL communicator (libnccl is dlopen-ed on first use)
           "wd_host.cc"]         # host-only: page-locking caller memory
HEADERS = ["wd_common.cuh", "wd_scan.cuh", "wd_seq.cuh", "wd_pack.cuh", "wd_kernels23.cuh", os.path.join("..", "..", "include", "welldup.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), tag=None):
    """``defines`` / ``tag``: a measurement variant (e.g. -DWD_PLANE_LD_L2_64B=0) built beside the product as
    libwelldup_<tag>.so, for A/B runs through ``bench.py --library``."""
    lib = LIB if tag is None else os.path.join(HERE, "libwelldup_%s.so" % tag)
    if tag is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ("" if tag is None else "." + tag) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    # --no-undefined: a kernel flavour that no translation unit instantiates must fail here, not at dlopen
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                 "-Xlinker", "--no-undefined", "-ldl"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    tags = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, tag=tags[0] if tags else None))
