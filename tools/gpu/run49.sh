#!/bin/bash
# CUDA-graph replay of the optimiser step's segments: parity + A/B
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r49_tests.txt 2>&1; tail -4 gpurun_out/r49_tests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
TRXL_TRAIN_GRAPHS=1 $B > gpurun_out/r49_graphs.json 2> gpurun_out/r49_graphs.err
$B > gpurun_out/r49_eager.json 2>/dev/null
python - <<'PY'
import json
for f in ("r49_graphs","r49_eager"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), r['launches_timed'], d['gpu_launches'], d['last_stats'][:3])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
tail -3 gpurun_out/r49_graphs.err
