#!/bin/bash
# 32 vs 64 targets per CTA on ONE box: lane (resident + zero-copy e2e) and the 704-tile CBCL lane, interleaved twice
L64="--library well_duplicates_b200/libwelldup_tpb64.so"
for round in 1 2; do
 for v in 32 64; do
  if [ $v = 64 ]; then L=$L64; else L=""; fi
  python bench.py --steps 40 --no-files --no-cpu-baseline --no-inflate $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lane tpb $v ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],2), 'logged', round(d['e2e_logged']['ms_per_step'],2))"
  python bench.py --config cbcl --steps 5 --no-cpu-baseline $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cbcl tpb $v ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],1))"
 done
done
