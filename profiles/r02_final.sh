#!/bin/bash
# the round's last call: parity, the bench line (with a wd_set_tuning sweep of the targets per CTA), the reference arm, the CBCL lane, ncu of the resident launch (traffic.json)
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests.log 2>&1; tail -2 $O/r02_gpu_tests.log; grep -E '^(FAILED|ERROR|E  )' $O/r02_gpu_tests.log | head -20 | cut -c1-250
S="targets_per_cta=32;targets_per_cta=64;targets_per_cta=128;targets_per_cta=256"
python bench.py --sweep-steps "$S" > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
python bench.py --config cbcl --steps 5 --sweep-steps "targets_per_cta=64;targets_per_cta=128;targets_per_cta=256" > $O/r02_bench_cbcl.json 2> $O/r02_bench_cbcl.err
bash profiles/r02_ncu_only.sh > $O/r02_ncu_only.log 2>&1; grep -E "sha|duration_us" $O/r02_traffic.json
python - <<PY
import json
d=json.load(open("$O/r02_bench_n1.json")); r=d["roofline"]
print("ms", d["ms_per_step"], "value", d["value"], "frac", r["frac"], "req", r["request_bound"]["achieved_g_lines_per_s"], "e2e", d["e2e"]["ms_per_step"], "logged", d["e2e_logged"]["ms_per_step"], "files", d["e2e_files"]["seconds"], d["counters_match_oracle"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for k, v in d["sweep_steps"].items(): print("   ", k, {a: round(b, 4) for a, b in v.items()})
c=json.load(open("$O/r02_bench_cbcl.json")); print("cbcl", c["ms_per_step"], c["e2e"]["ms_per_step"], c["counters_match_oracle"])
for k, v in c["sweep_steps"].items(): print("   ", k, {a: round(b, 4) for a, b in v.items()})
PY
