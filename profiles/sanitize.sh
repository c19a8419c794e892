#!/bin/bash
# compute-sanitizer over the small end-to-end invocation (stage 1 -> 2 -> 3, fused and two-pass kernels,
# exhaustive mode) and the golden-run parity tests:  gpurun -- bash profiles/sanitize.sh
O=gpurun_out
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; tail -3 $O/sanitize_${tool}_smoke.log
done
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "count_cli or get_seqs or prepare_cli or exhaustive_matches or exhaustive_errors or k3_filter" > $O/sanitize_memcheck_tests.log 2>&1
echo "memcheck tests rc=$?"; tail -4 $O/sanitize_memcheck_tests.log
