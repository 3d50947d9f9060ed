#!/bin/bash
# after factoring the host-side grouping into pure functions: full GPU suite + device-resident bench
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r52_tests.txt 2>&1; tail -3 gpurun_out/r52_tests.txt
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r52_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r52_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), d['last_stats'][:3])
PY
