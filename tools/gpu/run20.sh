cd $GRAFT_REPO_ROOT
for wl in c4_minigrid_gtrxl_synthetic c5_mortar_synthetic; do
timeout 1500 python bench.py --workload $wl --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; tail -3 gpurun_out/r2_bench_$wl.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_$wl.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['breakdown_s_per_update'], d['roofline']['kernel'][:50], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
except Exception as e: print("parse failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
