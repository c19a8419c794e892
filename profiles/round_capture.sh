#!/bin/bash
# Everything the round's numbers come from, in one call on the GPU box:
#   gpurun --timeout 2400 -- bash profiles/round_capture.sh
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
bash profiles/capture.sh > $O/capture.log 2>&1
python profiles/measure_configs.py --configs 4,3,5 > $O/configs.jsonl 2> $O/configs.err
bash profiles/capture_exhaustive.sh > $O/capture_exh.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
# a lane from .filter/.bcl.gz files on disk to counters (staging pipeline)
python profiles/measure_files_e2e.py --tiles 32 --distinct 8 > $O/files_e2e.json 2> $O/files_e2e.err; rm -rf /tmp/wd_run
tail -c 1500 $O/bench_n1.json; echo; cat $O/configs.jsonl | cut -c1-600; tail -2 $O/bench_ref.json | cut -c1-400; cut -c1-700 $O/files_e2e.json
