#!/bin/bash
# Targets per CTA of the fused kernel: one variant library per value, built before the call with
#   for t in 16 24 32 48 96; do python -m well_duplicates_b200.build --tag=tpb$t -DWD_FUSED_TPB=$t; done
# (the variants are not kept in the tree; the default library is the value wd_kernels23.cuh names)
B="python bench.py --steps 40 --e2e-steps 0 --no-files --no-cpu-baseline --no-inflate"
for t in ${@:-default 16 24 32 48 96 default}; do
  if [ $t = default ]; then L=""; else L="--library well_duplicates_b200/libwelldup_tpb$t.so"; fi
  $B $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tpb $t ms', round(d['ms_per_step'],4))"
done
