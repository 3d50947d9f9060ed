#!/bin/bash
# current build: tests, e2e rollout trace with 6 groups, bench
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r37_tests.txt 2>&1; tail -3 gpurun_out/r37_tests.txt
TRXL_E2E_TRACE=1 python tools/e2e_probe.py --rollouts 4 > gpurun_out/r37_e2e_trace.txt 2>&1; grep "trxl\] rollout trace\|probe\] rollout" gpurun_out/r37_e2e_trace.txt | tail -4
TRXL_E2E_TRACE=1 python tools/e2e_probe.py --rollouts 4 --profile > gpurun_out/r37_e2e_profile.txt 2>&1; grep -A25 "Name" gpurun_out/r37_e2e_profile.txt | cut -c1-80,150-260 | head -30
python bench.py --steps 3 --warmup 3 > gpurun_out/r37_bench.json 2> gpurun_out/r37_bench.err; tail -c 150 gpurun_out/r37_bench.json
