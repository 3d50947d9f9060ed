cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_tests9.log; cat gpurun_out/r2_tests9.log
TRXL_E2E_TRACE=1 timeout 300 python tools/e2e_probe.py --rollouts 4 2>&1 | tail -4
for wl in c3_minigrid_synthetic c1_poc_synthetic c2_cartpole_synthetic; do
timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; tail -2 gpurun_out/r2_bench_$wl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['rollout_s_per_update'], 'cpu', d['cpu_baseline']['value'], d['roofline']['kernel'][:40], d['roofline']['frac'])
PY
done
