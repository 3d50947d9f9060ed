#!/bin/bash
# one-launch column reductions: parity + A/B
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r46_tests.txt 2>&1; tail -3 gpurun_out/r46_tests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r46_onepass.json 2>/dev/null
TRXL_EW_ONEPASS=0 $B > gpurun_out/r46_twopass.json 2>/dev/null
python - <<'PY'
import json
for f in ("r46_onepass","r46_twopass"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3))
PY
