#!/bin/bash
# ncu captures behind profiles/r01_*: run on the GPU box (gpurun -- bash profiles/capture.sh).
# Reports are summarised on the box; only the text/CSV summaries travel back (gpurun_out/ is capped at 64 MiB).
set -x
O=gpurun_out
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
# e2e step (planes in pinned host memory, 16 tile groups per step): launches 0-4 are resident,
# 5-20 the first mapped step, 21-36 the timed one
ncu --set full --metrics pcie__read_bytes.sum,pcie__write_bytes.sum \
    --clock-control none -k regex:fused_count -c 16 -s 21 -o /tmp/fused_zc -f \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --e2e-mode zerocopy --no-cpu-baseline > $O/ncu_zc.log 2>&1
python profiles/summarize_ncu.py /tmp/fused_zc.ncu-rep > $O/fused_zero_copy_ncu.txt
ncu -i /tmp/fused_zc.ncu-rep --page raw --csv > $O/fused_zero_copy_raw.csv
# resident launch
ncu --set full --clock-control none --import-source on -k regex:fused_count -c 1 -s 4 -o /tmp/fused_res -f \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $O/ncu_res.log 2>&1
python profiles/summarize_ncu.py /tmp/fused_res.ncu-rep > $O/fused_resident_ncu.txt
ncu -i /tmp/fused_res.ncu-rep --page raw --csv > $O/fused_resident_raw.csv
ncu -i /tmp/fused_res.ncu-rep --page source --csv --print-source sass > $O/fused_resident_sass.csv
# launch list of the bench command (cold, serialised): the kernel's share of the step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/ncu_l.log 2>&1
python profiles/make_traffic.py $O/fused_resident_raw.csv $O/fused_zero_copy_raw.csv > $O/traffic.json
ls -la $O
