#!/bin/bash
# fused softmax / dscore row epilogues; 6 worker groups
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r35_tests.txt 2>&1; tail -3 gpurun_out/r35_tests.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r35_bench.json 2> gpurun_out/r35_bench.err; tail -c 200 gpurun_out/r35_bench.json
TRXL_ATTN_FUSED_ROWS=0 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r35_bench_unfused.json 2> gpurun_out/r35_bench_unfused.err; tail -c 100 gpurun_out/r35_bench_unfused.json
