O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_count -c 1 -s 4 -o /tmp/fused_res -f \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-inflate > $O/ncu_res.log 2>&1
python profiles/summarize_ncu.py /tmp/fused_res.ncu-rep > $O/fused_resident_ncu.txt
ncu -i /tmp/fused_res.ncu-rep --page raw --csv > $O/fused_resident_raw.csv
ncu -i /tmp/fused_res.ncu-rep --page source --csv --print-source sass > $O/fused_resident_sass.csv
grep -E "duration|dram__bytes_read|inst_executed.sum|issue_active|pipe_alu|no_instruction|long_score|registers" $O/fused_resident_ncu.txt
