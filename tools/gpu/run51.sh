#!/bin/bash
# final build (training graphs on): c4 / c5 device-resident
cd /root/repo
for w in c4_minigrid_gtrxl_synthetic c5_mortar_synthetic; do
  timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r51_$w.json 2>> gpurun_out/r51.err
done
python - <<'PY'
import json
for w in ("c4_minigrid_gtrxl_synthetic","c5_mortar_synthetic"):
    try:
        d=json.loads(open('gpurun_out/r51_%s.json'%w).read().strip().splitlines()[-1]); r=d['roofline']
        print(w, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], r.get('frac'), d['last_stats'][:3])
    except Exception as ex:
        print(w, 'FAILED', ex)
PY
tail -3 gpurun_out/r51.err
