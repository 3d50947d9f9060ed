#!/bin/bash
# final c3 bench (all arms) + kineto + attention ncu refresh
cd /root/repo
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r43_bench.err; tail -c 120 gpurun_out/r2_bench_c3_final.json
python tools/profile_step.py --kineto gpurun_out/kernels_c3.txt > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tma_gemm -s 4 -c 2 -o gpurun_out/attn_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu43b.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_step.py --range --rollout-steps 2 --minibatches 2 > gpurun_out/ncu43a.log 2>&1
TRXL_TC_BN=64 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r43_bn64.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r43_bnauto.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_bench_c3_final","r43_bn64","r43_bnauto"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}
    print(f, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), round(r['avg_launch_ms']*1e3,1), round(r['bwd']['avg_launch_ms']*1e3,1), e.get('value'), e.get('rollout_s_per_update'), e.get('train_s_per_update'))
PY
