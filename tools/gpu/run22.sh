cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
run() { # name, args...
  name=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 "$@" > gpurun_out/r2_bench8_$name.json 2> gpurun_out/r2_bench8_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench8_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d[k] for k in ('metric','value','ms_per_step','n_gpus','scaling')}, d['breakdown_s_per_update'], 'e2e', (d['e2e'] or {}).get('value'), (d['e2e'] or {}).get('rollout_s_per_update'), (d['e2e'] or {}).get('train_s_per_update'))
except Exception as e:
    print('$name parse failed', e); print(open('gpurun_out/r2_bench8_$name.err').read()[-1500:])
PY
}
run c3_weak --steps 3 --warmup 3
run c4_strong --workload c4_minigrid_gtrxl_synthetic --scaling strong --steps 2 --warmup 2
run c5_strong --workload c5_mortar_synthetic --scaling strong --steps 1 --warmup 1 --no-e2e
run c3_strong --scaling strong --steps 3 --warmup 3 --no-e2e
