#!/bin/bash
# Round 2: everything the round's numbers come from, one GPU (gpurun --timeout 2400 -- bash profiles/r02_capture.sh)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests.log 2>&1; tail -3 $O/r02_gpu_tests.log
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
for c in stage1 exhaustive cbcl; do
  python bench.py --config $c --steps 5 > $O/r02_bench_$c.json 2> $O/r02_bench_$c.err
  python bench.py --config $c --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref_$c.json 2>> $O/r02_bench_$c.err
done
# ncu: the resident launch of the fused kernel (the same command has just exited 0 without ncu)
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-inflate"
$CMD > $O/r02_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_count -c 1 -s 4 -o /tmp/fused_res -f $CMD > $O/r02_ncu_res.log 2>&1
python profiles/summarize_ncu.py /tmp/fused_res.ncu-rep > $O/r02_fused_resident_ncu.txt
ncu -i /tmp/fused_res.ncu-rep --page raw --csv > $O/r02_fused_resident_raw.csv
ncu -i /tmp/fused_res.ncu-rep --page source --csv --print-source sass > $O/r02_fused_resident_sass.csv
python profiles/make_traffic.py $O/r02_fused_resident_raw.csv > $O/r02_traffic.json
# launch list of the bench command (cold, serialised): the kernel's share of the step
CMD2="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-inflate"
$CMD2 > $O/r02_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $CMD2 > $O/r02_ncu_l.log 2>&1
tail -c 1200 $O/r02_bench_n1.json; echo
for c in stage1 exhaustive cbcl; do cut -c1-500 $O/r02_bench_$c.json; echo; done
grep -E "duration|dram__bytes_read|inst_executed.sum|issue_active|pipe_alu|long_score|registers|warps_active" $O/r02_fused_resident_ncu.txt
