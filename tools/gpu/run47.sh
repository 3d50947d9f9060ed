#!/bin/bash
# final build: c4 / c5 device-resident benches, c1 / c2 full
cd /root/repo
for w in c4_minigrid_gtrxl_synthetic c5_mortar_synthetic; do
  timeout 500 python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r47_$w.json 2>> gpurun_out/r47.err
done
for w in c1_poc_synthetic c2_cartpole_synthetic; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r47_$w.json 2>> gpurun_out/r47.err
done
python - <<'PY'
import json
for w in ("c4_minigrid_gtrxl_synthetic","c5_mortar_synthetic","c1_poc_synthetic","c2_cartpole_synthetic"):
    try:
        d=json.loads(open('gpurun_out/r47_%s.json'%w).read().strip().splitlines()[-1]); r=d['roofline']; e=d.get('e2e') or {}
        print(w, round(d['value']), round(d['ms_per_step'],1), d['breakdown_s_per_update'], r.get('frac'), e.get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as ex:
        print(w, 'FAILED', ex)
PY
tail -3 gpurun_out/r47.err
