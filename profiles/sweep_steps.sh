# early-exit schedule sweep of the fused kernel: WELLDUP_STEPS=first,next cycles per round
for sch in 8,4 4,4 8,2 4,2 2,2; do WELLDUP_STEPS=$sch python bench.py --steps 10 --e2e-steps 0 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('lev', '$sch', 'ms', round(d['ms_per_step'],4))"; done
for sch in 4,4 4,2 2,2; do WELLDUP_STEPS=$sch python bench.py --steps 10 --e2e-steps 0 --no-cpu-baseline --hamming 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('ham', '$sch', 'ms', round(d['ms_per_step'],4))"; done
