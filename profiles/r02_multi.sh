#!/bin/bash
# Round 2, N GPUs of one box (gpurun --gpus 2 -- bash profiles/r02_multi.sh 2): the flowcell driver against the
# reference's reports, then the bench line at N (library NCCL all-reduce, parity flag, files -> counters)
N=${1:-2}
O=gpurun_out
bash tests/multi_gpu_check.sh $N > $O/r02_mgc_n$N.log 2>&1; grep identical $O/r02_mgc_n$N.log; tail -2 $O/r02_mgc_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N ${BENCH_ARGS:-} > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("$O/r02_bench_n$N.json"))
    print("N=$N ms", round(d["ms_per_step"], 4), "value", d["value"], "parity", d.get("counters_match_oracle"),
          "e2e ms", d["e2e"]["ms_per_step"], "logged", d.get("e2e_logged", {}).get("ms_per_step"),
          "files s", d.get("e2e_files", {}).get("seconds"), "flowcell s", d.get("e2e_files", {}).get("flowcell_seconds_at_this_rate"),
          "clocks", d["clocks"])
    if "sweep_steps" in d: print(json.dumps(d["sweep_steps"]))
except Exception as e:
    print("bench failed:", e)
PY
tail -5 $O/r02_bench_n$N.err
if [ "$N" = "2" ]; then
  python bench.py --e2e-steps 0 --no-files --no-cpu-baseline --no-inflate > $O/r02_bench_quick.json 2> $O/r02_bench_quick.err
  python -c "import json; d=json.load(open('$O/r02_bench_quick.json')); print('N=1 quick ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])"
fi
