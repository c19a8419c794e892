#!/bin/bash
# e2e (zero-copy over PCIe) with 128-byte L2 prefetch size on the plane loads vs plain loads
O=gpurun_out
B="python bench.py --e2e-steps 3 --e2e-mode zerocopy --no-files --no-cpu-baseline --no-inflate"
$B > $O/r02_ab3_plain.json 2> $O/r02_ab3_plain.err
$B --library well_duplicates_b200/libwelldup_l2_128.so > $O/r02_ab3_l2_128.json 2> $O/r02_ab3_l2_128.err
$B --sweep-steps "head_planes=1;head_planes=3;2,1;2,2;2,4;step0=2 step1=3" > $O/r02_ab3_sweep.json 2> $O/r02_ab3_sweep.err
for f in plain l2_128 sweep; do python - <<PY
import json
try:
    d=json.load(open("$O/r02_ab3_$f.json"))
    print("$f", "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],2), "logged", round(d["e2e_logged"]["ms_per_step"],2), "pull req/s", d["e2e"]["pull_requests_per_s_per_gpu"], "dma", d["e2e"]["h2d_dma_gb_per_s"], d["e2e"]["head_planes_by_dma"])
    if "sweep_steps" in d: print(json.dumps(d["sweep_steps"]))
except Exception as e:
    print("$f failed", e)
PY
done
