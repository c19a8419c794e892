#!/bin/bash
# ncu of the resident launch of the fused kernel only -> profiles/traffic.json (gpurun -- bash profiles/r02_ncu_only.sh)
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-files --no-cpu-baseline --no-inflate"
$CMD > $O/r02_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_count -c 1 -s 4 -o /tmp/fused_res -f $CMD > $O/r02_ncu_res.log 2>&1
python profiles/summarize_ncu.py /tmp/fused_res.ncu-rep > $O/r02_fused_resident_ncu.txt
ncu -i /tmp/fused_res.ncu-rep --page raw --csv > $O/r02_fused_resident_raw.csv
ncu -i /tmp/fused_res.ncu-rep --page source --csv --print-source sass > $O/r02_fused_resident_sass.csv
python profiles/make_traffic.py $O/r02_fused_resident_raw.csv > $O/r02_traffic.json
cat $O/r02_traffic.json
grep -E "duration|dram__bytes_read|inst_executed.sum|issue_active|long_score|registers" $O/r02_fused_resident_ncu.txt
