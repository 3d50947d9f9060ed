cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; tail -2 gpurun_out/r2_bench9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench9.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['breakdown_s_per_update']); r=d['roofline']; print(r['frac'], r['avg_launch_ms'], r['bwd'])
PY
