cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_tests6.log; cat gpurun_out/r2_tests6.log
timeout 600 python tools/profile_step.py --kineto gpurun_out/r2_kineto_update_ga.txt > /dev/null 2>&1
cut -c1-92,196-260 gpurun_out/r2_kineto_update_ga.txt | head -24
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['breakdown_s_per_update'])
PY
