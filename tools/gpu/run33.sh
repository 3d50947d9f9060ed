#!/bin/bash
# conv gather: cluster split-K + coalesced epilogue; rollout groups sweep
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r33_tests.txt 2>&1; tail -5 gpurun_out/r33_tests.txt
python tools/rf_trace.py --steps 1 > gpurun_out/r33_trace.txt 2>&1; grep -c trace gpurun_out/r33_trace.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r33_bench.json 2> gpurun_out/r33_bench.err; tail -c 300 gpurun_out/r33_bench.json
for g in 3 4; do
  TRXL_ROLLOUT_GROUPS=$g python tools/e2e_probe.py > gpurun_out/r33_e2e_groups$g.txt 2>&1; tail -3 gpurun_out/r33_e2e_groups$g.txt
done
