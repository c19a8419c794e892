#!/bin/bash
# ncu captures of exhaustive mode (gpurun -- bash profiles/capture_exhaustive.sh)
O=gpurun_out
python profiles/run_exhaustive.py 3 > $O/exh_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/exh_launches.csv \
    python profiles/run_exhaustive.py 2 > $O/exh_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:exh_|dense_pack' -c 5 -o /tmp/exh -f \
    python profiles/run_exhaustive.py 1 > $O/exh_ncu.log 2>&1
python profiles/summarize_ncu.py /tmp/exh.ncu-rep > $O/exhaustive_v3_ncu.txt
ncu -i /tmp/exh.ncu-rep --page source --csv --print-source sass -k regex:exh_compare > $O/exh_compare_sass.csv 2>/dev/null
cat $O/exh_plain.log
