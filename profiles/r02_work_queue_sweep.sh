#!/bin/bash
# targets per grab of the work queues, CTAs per launch and head-plane groups through wd_set_tuning, in one process per config: lane (resident + zero-copy), CBCL 704 and 96 tiles
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests.log 2>&1; tail -15 $O/r02_gpu_tests.log | cut -c1-300; grep -q passed $O/r02_gpu_tests.log && ! grep -q failed $O/r02_gpu_tests.log || exit 1
S="targets_per_grab=1;targets_per_grab=2;targets_per_grab=3;targets_per_grab=4;grid_ctas=74;grid_ctas=148;grid_ctas=296;grid_ctas=444;grid_ctas=592;grid_ctas=888;grid_ctas=148 targets_per_grab=2;grid_ctas=296 targets_per_grab=2;head_groups=4;head_groups=8;head_groups=32;head_groups=8 grid_ctas=296;head_groups=4 grid_ctas=296;head_groups=8 grid_ctas=148"
C="targets_per_grab=1;targets_per_grab=2;targets_per_grab=4;grid_ctas=592;grid_ctas=296"
python bench.py --steps 30 --no-files --no-cpu-baseline --no-inflate --sweep-steps "$S" > $O/r02_grab_lane.json 2>/dev/null
python bench.py --config cbcl --steps 5 --no-cpu-baseline --sweep-steps "$C" > $O/r02_grab_cbcl704.json 2>/dev/null
python bench.py --config cbcl --cbcl-tiles 96 --steps 20 --no-cpu-baseline --sweep-steps "$C" > $O/r02_grab_cbcl96.json 2>/dev/null
python - <<PY
import json
for n in ("lane", "cbcl704", "cbcl96"):
    d = json.load(open("$O/r02_grab_%s.json" % n))
    print(n, "default", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 2))
    for k, v in d["sweep_steps"].items():
        print("   ", k, {a: round(b, 4) for a, b in v.items()})
PY
