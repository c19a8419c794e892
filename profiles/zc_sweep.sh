# zero-copy e2e (planes in pinned host memory) vs early-exit schedule
for sch in 8,4 6,2 5,1 4,2 4,1 3,1; do WELLDUP_STEPS=$sch python bench.py --e2e-mode zerocopy --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('zc lev', '$sch', 'ms', round(d['e2e_zero_copy']['ms_per_step'],2))"; done
