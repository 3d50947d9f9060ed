#!/bin/bash
# 2 GPUs: NCCL gradient test + bench
cd /root/repo
python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu -k two_gpu > gpurun_out/r39_tests_2gpu.txt 2>&1; tail -3 gpurun_out/r39_tests_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r39_bench_2gpu.json 2> gpurun_out/r39_bench_2gpu.err; tail -c 300 gpurun_out/r39_bench_2gpu.json
