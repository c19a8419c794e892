#!/bin/bash
# parity + timing of the fused kernel after a change (gpurun -- bash profiles/ab_fused.sh)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --e2e-steps 0 --no-cpu-baseline --no-inflate 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lev ms',d['ms_per_step'],'targets/s',d['value'])"
python bench.py --hamming --steps 20 --e2e-steps 0 --no-cpu-baseline --no-inflate 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ham ms',d['ms_per_step'])"
python bench.py --steps 10 --e2e-steps 0 --no-cpu-baseline --no-inflate --sweep-steps "${SWEEP:-8,4;8,2;6,2;4,2;4,4}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['sweep_steps']))"
