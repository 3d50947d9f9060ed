cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests3.log; cat gpurun_out/r2_tests3.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['breakdown_s_per_update']); print(d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/r2_bench2.err
