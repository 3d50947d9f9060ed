#!/bin/bash
# training graphs on by default: full tests + full bench
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r50_tests.txt 2>&1; tail -4 gpurun_out/r50_tests.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r50_bench.err; tail -c 120 gpurun_out/r2_bench_c3_final.json
python tools/profile_step.py --kineto gpurun_out/kernels_c3.txt > /dev/null 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c3_final.json').read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}
print(round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), round(r['avg_launch_ms']*1e3,1), round(r['bwd']['avg_launch_ms']*1e3,1), r['launches_timed'], e.get('value'), e.get('rollout_s_per_update'), e.get('train_s_per_update'), d['cpu_baseline']['value'], d['gpu_launches'])
PY
