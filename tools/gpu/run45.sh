#!/bin/bash
# final: tests, c3 bench (all arms), kineto, launch list, conv ncu
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r45_tests.txt 2>&1; tail -3 gpurun_out/r45_tests.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r45_bench.err; tail -c 120 gpurun_out/r2_bench_c3_final.json
python tools/profile_step.py --kineto gpurun_out/kernels_c3.txt > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_step.py --range --rollout-steps 2 --minibatches 2 > gpurun_out/ncu45a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tc_conv -c 8 -o gpurun_out/tcconv_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu45e.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c3_final.json').read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}
print(round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), round(r['avg_launch_ms']*1e3,1), round(r['bwd']['avg_launch_ms']*1e3,1), e.get('value'), e.get('rollout_s_per_update'), e.get('train_s_per_update'), d['cpu_baseline']['value'])
PY
grep "tc_conv" gpurun_out/kernels_c3.txt | cut -c1-90,190-260
