#!/bin/bash
# wd_set_tuning sweeps of the targets per CTA and of a cap on the resident CTAs per SM, in one process per config
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "random_configurations or zero_copy or mapped" 2>&1 | tail -2
S="targets_per_cta=32;targets_per_cta=64;targets_per_cta=128;targets_per_cta=192;targets_per_cta=256;ctas_per_sm=1 targets_per_cta=32;ctas_per_sm=1 targets_per_cta=64;ctas_per_sm=2 targets_per_cta=32;ctas_per_sm=2 targets_per_cta=64;ctas_per_sm=1 targets_per_cta=128;ctas_per_sm=4 targets_per_cta=32;ctas_per_sm=4 targets_per_cta=64"
python bench.py --steps 30 --no-files --no-cpu-baseline --no-inflate --sweep-steps "$S" > $O/r02_tpb_lane.json 2>/dev/null
python bench.py --config cbcl --steps 5 --no-cpu-baseline --sweep-steps "targets_per_cta=64;targets_per_cta=128;targets_per_cta=256;ctas_per_sm=2 targets_per_cta=128;ctas_per_sm=1 targets_per_cta=64;ctas_per_sm=4 targets_per_cta=64" > $O/r02_tpb_cbcl704.json 2>/dev/null
python - <<PY
import json
for n in ("lane", "cbcl704"):
    d = json.load(open("$O/r02_tpb_%s.json" % n))
    print(n, "default", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 2))
    for k, v in d["sweep_steps"].items():
        print("   ", k, {a: round(b, 4) for a, b in v.items()})
PY
