#!/bin/bash
# 4 worker groups by default; conv epilogue details; fresh kineto of one update
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r34_tests.txt 2>&1; tail -3 gpurun_out/r34_tests.txt
python tools/rf_trace.py --steps 1 > gpurun_out/r34_trace.txt 2>&1; grep "conv-trace\] BN" gpurun_out/r34_trace.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r34_bench.json 2> gpurun_out/r34_bench.err; tail -c 200 gpurun_out/r34_bench.json
python tools/profile_step.py --kineto gpurun_out/r34_kineto.txt > /dev/null 2>&1
for g in 6 8; do
  TRXL_ROLLOUT_GROUPS=$g python tools/e2e_probe.py > gpurun_out/r34_e2e_groups$g.txt 2>&1; tail -2 gpurun_out/r34_e2e_groups$g.txt
done
