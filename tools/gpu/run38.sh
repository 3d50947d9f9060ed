#!/bin/bash
# sequence-tagged actions, parallel wgrad reduce
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r38_tests.txt 2>&1; tail -3 gpurun_out/r38_tests.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r38_bench.json 2> gpurun_out/r38_bench.err; tail -c 150 gpurun_out/r38_bench.json
python tools/e2e_probe.py --rollouts 5 > gpurun_out/r38_e2e.txt 2>&1; tail -3 gpurun_out/r38_e2e.txt
python tools/e2e_probe.py --rollouts 4 --profile > gpurun_out/r38_e2e_profile.txt 2>&1; grep -A12 "Name" gpurun_out/r38_e2e_profile.txt | cut -c1-80,150-260 | head -14
