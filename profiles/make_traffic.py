#!/usr/bin/env python3
"""ncu raw CSV pages -> profiles/traffic.json (per-launch / per-step byte counts bench.py quotes).

    python profiles/make_traffic.py RESIDENT_RAW.csv [ZERO_COPY_RAW.csv] > profiles/traffic.json

RESIDENT: one launch of fused_count_kernel over 96 tiles with planes in HBM.
ZERO_COPY: the launches of one e2e step (planes in pinned host memory; one launch per tile group).
The file carries the hash of the kernel sources it was captured from (bench.kernel_source_hash): bench.py
quotes these numbers only while the sources it runs are the same."""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    return hdr, data


def col(hdr, data, name):
    i = hdr.index(name)
    return [float(r[i].replace(",", "")) for r in data]


def unit_scale(hdr, path, name):
    rows = list(csv.reader(open(path)))
    u = rows[1][hdr.index(name)].lower()
    return {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "sector": 1, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)


def main(res, zc=None):
    from bench import kernel_source_hash
    out = {"kernel_source_sha256": kernel_source_hash()}
    h, d = load(res)
    rd = col(h, d, "dram__bytes_read.sum")[0] * unit_scale(h, res, "dram__bytes_read.sum")
    wr = col(h, d, "dram__bytes_write.sum")[0] * unit_scale(h, res, "dram__bytes_write.sum")
    out["fused"] = int(rd + wr)
    out["fused_detail"] = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                           "duration_us": col(h, d, "gpu__time_duration.sum")[0] * unit_scale(h, res, "gpu__time_duration.sum"),
                           "l1_requested_sectors": int(col(h, d, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")[0])}
    if zc is None:
        json.dump(out, sys.stdout, indent=1)
        print()
        return
    h, d = load(zc)
    pr = sum(col(h, d, "pcie__read_bytes.sum")) * unit_scale(h, zc, "pcie__read_bytes.sum")
    sec = sum(col(h, d, "syslts__t_sectors_srcunit_tex_aperture_sysmem_op_read_lookup_miss.sum"))
    out["zero_copy_pcie_read_bytes"] = int(pr)
    out["zero_copy_sysmem_read_bytes"] = int(sec * 32)
    out["zero_copy_detail"] = {"launches": len(d), "sysmem_sectors_read": int(sec),
                               "dram_bytes_read": int(sum(col(h, d, "dram__bytes_read.sum")) * unit_scale(h, zc, "dram__bytes_read.sum")),
                               "kernel_us_serialised": sum(col(h, d, "gpu__time_duration.sum")) * unit_scale(h, zc, "gpu__time_duration.sum")}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(*sys.argv[1:3])
