cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_tests7.log; cat gpurun_out/r2_tests7.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; tail -3 gpurun_out/r2_bench7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches','scaling')}); print(d['breakdown_s_per_update']); print(json.dumps(d['roofline'],indent=1)); print(d['e2e']); print(d['cpu_baseline'])
PY
for wl in c1_poc_synthetic c2_cartpole_synthetic; do
timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; tail -2 gpurun_out/r2_bench_$wl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step')}, d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'][:40], d['roofline']['frac'])
PY
done
