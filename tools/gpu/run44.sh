#!/bin/bash
# merged dgrad2 (four parity classes in one 128-wide GEMM): parity + A/B
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r44_tests.txt 2>&1; tail -3 gpurun_out/r44_tests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r44_merged.json 2>/dev/null
TRXL_CONV_DGRAD2_CLASSES=1 $B > gpurun_out/r44_classes.json 2>/dev/null
python - <<'PY'
import json
for f in ("r44_merged","r44_classes"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3))
PY
python tools/profile_step.py --kineto gpurun_out/r44_kineto.txt > /dev/null 2>&1; grep "tc_conv" gpurun_out/r44_kineto.txt | cut -c1-90,190-260
