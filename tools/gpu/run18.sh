cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or c3_dims" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['breakdown_s_per_update'])
PY
TRXL_E2E_TRACE=1 timeout 300 python tools/e2e_probe.py --rollouts 4 2>&1 | tail -3
