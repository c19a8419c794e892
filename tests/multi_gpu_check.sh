#!/bin/bash
# Multi-GPU check (run on a box with >= 2 B200s:  gpurun --gpus 2 -- bash tests/multi_gpu_check.sh):
# the flowcell driver under torchrun must print exactly what the reference printed for the same run.
set -e
N=${1:-2}
G=tests/golden
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        -m well_duplicates_b200.flowcell -f $G/locs/hex_small_n40_s13.list -r $G/$1 -q "${@:2}"; }
run run_bcl -s hiseq_x -i 1,2 -t 1101 -l 5 --cycles 0-14 > /tmp/two_lanes.out
cmp /tmp/two_lanes.out $G/count/two_lanes.stdout && echo "two_lanes: identical on $N GPUs"
run run_bcl -s hiseq_x -i 1 -t 1101,1102 -l 5 --cycles 0-14 > /tmp/lev_default.out
cmp /tmp/lev_default.out $G/count/lev_default.stdout && echo "lev_default: identical on $N GPUs"
run run_cbcl -s 2488 -i 1 -t 1101,2101 -l 5 --cycles 0-14 > /tmp/cbcl_default.out
cmp /tmp/cbcl_default.out $G/count/cbcl_default.stdout && echo "cbcl_default: identical on $N GPUs"
# without -q every rank logs its own tiles (duplicate pairs incl.): stdout must not change, and every pair of the
# reference's log must appear exactly once across the ranks' stderr
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    -m well_duplicates_b200.flowcell -f $G/locs/hex_small_n40_s13.list -r $G/run_bcl -s hiseq_x -i 1,2 -t 1101 -l 5 --cycles 0-14 \
    > /tmp/two_lanes_logged.out 2> /tmp/two_lanes_logged.err
cmp /tmp/two_lanes_logged.out $G/count/two_lanes.stdout && echo "two_lanes (logged): identical on $N GPUs"
[ "$(grep -c '^edit distance' /tmp/two_lanes_logged.err)" = "$(grep -c '^edit distance' $G/count/two_lanes.stderr)" ] && \
    [ "$(grep '^well seq\|^center seq\|^edit' /tmp/two_lanes_logged.err | sort | md5sum)" = "$(grep '^well seq\|^center seq\|^edit' $G/count/two_lanes.stderr | sort | md5sum)" ] && \
    echo "two_lanes (logged): the ranks' logs hold the reference's duplicate pairs"
