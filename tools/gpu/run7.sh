cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_timed_path.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -30 > gpurun_out/r2_tests_2gpu.log; cat gpurun_out/r2_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e']); print(d['breakdown_s_per_update'])
PY
tail -5 gpurun_out/r2_bench_2gpu.err
