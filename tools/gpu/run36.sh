#!/bin/bash
# A/B on one box: fused row epilogues, cluster split-K
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r36_tests.txt 2>&1; tail -3 gpurun_out/r36_tests.txt
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/r36_A.json 2>/dev/null
TRXL_ATTN_FUSED_ROWS=0 $B > gpurun_out/r36_B.json 2>/dev/null
TRXL_ATTN_FUSED_ROWS=0 TRXL_TC_CLUSTER_SPLITK=0 $B > gpurun_out/r36_C.json 2>/dev/null
TRXL_TC_CLUSTER_SPLITK=0 $B > gpurun_out/r36_D.json 2>/dev/null
$B > gpurun_out/r36_A2.json 2>/dev/null
python - <<'PY'
import json
for f in "A B C D A2".split():
    d=json.loads(open('gpurun_out/r36_%s.json'%f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], round(r['frac'],3), round(r['avg_launch_ms']*1e3,1), round(r['bwd']['avg_launch_ms']*1e3,1))
PY
