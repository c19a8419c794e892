#!/bin/bash
# parity + timing after a kernel change (gpurun -- bash profiles/r02_quick.sh)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_quick_tests.log 2>&1; tail -3 $O/r02_quick_tests.log
python bench.py --e2e-steps 2 --e2e-mode zerocopy --no-files --no-cpu-baseline --no-inflate > $O/r02_quick.json 2> $O/r02_quick.err
python bench.py --hamming --e2e-steps 0 --no-files --no-cpu-baseline --no-inflate > $O/r02_quick_ham.json 2>> $O/r02_quick.err
python - <<PY
import json
for f in ("r02_quick", "r02_quick_ham"):
    try:
        d = json.load(open("$O/%s.json" % f))
        print(f, "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "lines/s", round(d["roofline"]["request_bound"]["achieved_g_lines_per_s"], 2),
              "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"], 2), "logged", d.get("e2e_logged") and round(d["e2e_logged"]["ms_per_step"], 2), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $O/r02_quick.err
