#!/bin/bash
# rollout: prefetched windows + ILP passes, 8-warp conv producers, rollout_store, flatten in conv3 epilogue
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r32_tests.txt 2>&1; tail -5 gpurun_out/r32_tests.txt
python tools/rf_trace.py --steps 1 > gpurun_out/r32_trace.txt 2>&1; grep -c trace gpurun_out/r32_trace.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r32_bench.json 2> gpurun_out/r32_bench.err; tail -c 300 gpurun_out/r32_bench.json
