cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; tail -2 gpurun_out/r2_bench_c3_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c3_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['breakdown_s_per_update']); r=d['roofline']; print(r['frac'], r['avg_launch_ms'], r['bwd'], r['traffic']); print(d['e2e']); print(d['cpu_baseline']['value'], d['cpu_baseline']['extrapolation']['factor'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tma_gemm -s 4 -c 2 -o gpurun_out/attn_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu29.log 2>&1; tail -1 gpurun_out/ncu29.log
timeout 600 python tools/profile_step.py --kineto gpurun_out/r2_kineto_update_final.txt > /dev/null 2>&1
cut -c1-92,196-260 gpurun_out/r2_kineto_update_final.txt | head -14
