#!/bin/bash
# window prefetch in the fused rollout kernel: parity, traces, bench
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r31_tests.txt 2>&1; tail -5 gpurun_out/r31_tests.txt
python tools/rf_trace.py --steps 1 > gpurun_out/r31_trace.txt 2>&1; grep -c rf-trace gpurun_out/r31_trace.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r31_bench.json 2> gpurun_out/r31_bench.err; tail -c 1500 gpurun_out/r31_bench.json
