#!/bin/bash
# round 2, call 2: GPU parity tests, then the plane-load flavour A/B (ld.global.nc.L2::64B vs plain __ldg)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests.log 2>&1; tail -5 $O/r02_gpu_tests.log
python bench.py --steps 20 --e2e-steps 2 --e2e-mode zerocopy --no-cpu-baseline --no-inflate > $O/r02_ab_l2_64.json 2> $O/r02_ab_l2_64.err
python bench.py --steps 20 --e2e-steps 2 --e2e-mode zerocopy --no-cpu-baseline --no-inflate --library well_duplicates_b200/libwelldup_plainld.so > $O/r02_ab_plain.json 2> $O/r02_ab_plain.err
python bench.py --hamming --steps 20 --e2e-steps 0 --no-cpu-baseline --no-inflate > $O/r02_ab_l2_64_ham.json 2>> $O/r02_ab_l2_64.err
for f in l2_64 plain l2_64_ham; do python - <<PY
import json
try:
    d=json.load(open("$O/r02_ab_$f.json"))
    print("$f", "ms", round(d["ms_per_step"],4), "frac", d.get("roofline",{}).get("frac"), "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"],2), "logged", d.get("e2e_logged") and round(d["e2e_logged"]["ms_per_step"],2), d.get("counters_match_oracle"))
except Exception as e:
    print("$f failed", e)
PY
done
tail -3 $O/r02_ab_l2_64.err
