#!/bin/bash
# 16-deep k-blocks (4-stage ring) for the 256-wide attention tiles
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r42_tests.txt 2>&1; tail -3 gpurun_out/r42_tests.txt
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r42_bench.json 2> gpurun_out/r42_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r42_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), round(r['avg_launch_ms']*1e3,1), round(r['bwd']['avg_launch_ms']*1e3,1))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
