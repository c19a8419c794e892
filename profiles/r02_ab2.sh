#!/bin/bash
# round 2, call 3: parity of the reworked fused kernel (exact centre reads, Eq table), occupancy variants, schedule sweeps
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests2.log 2>&1; tail -5 $O/r02_gpu_tests2.log
B="python bench.py --steps 20 --no-cpu-baseline --no-inflate"
$B --e2e-steps 2 --e2e-mode zerocopy --sweep-steps "8,4;8,2;6,2;7,3;8,4,16;8,4 visit_order=0" > $O/r02_ab2_default.json 2> $O/r02_ab2_default.err
$B --e2e-steps 0 --library well_duplicates_b200/libwelldup_occ8.so --sweep-steps "8,4;8,2;6,2" > $O/r02_ab2_occ8.json 2> $O/r02_ab2_occ8.err
$B --e2e-steps 0 --library well_duplicates_b200/libwelldup_occ5.so > $O/r02_ab2_occ5.json 2> $O/r02_ab2_occ5.err
$B --e2e-steps 0 --hamming > $O/r02_ab2_ham.json 2> $O/r02_ab2_ham.err
for f in default occ8 occ5 ham; do python - <<PY
import json
try:
    d=json.load(open("$O/r02_ab2_$f.json"))
    print("$f", "ms", round(d["ms_per_step"],4), "frac", d.get("roofline",{}).get("frac"), "lines", d.get("roofline",{}).get("needed",{}).get("plane_lines_128B"), "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],2), "logged", d.get("e2e_logged") and round(d["e2e_logged"]["ms_per_step"],2))
    print("   sweep", json.dumps(d.get("sweep_steps")))
except Exception as e:
    print("$f failed", e)
PY
done
tail -3 $O/r02_ab2_default.err
