#!/bin/bash
# final evidence, part 1: tests, benches of every configuration, kineto of one update
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r40_tests.txt 2>&1; tail -3 gpurun_out/r40_tests.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r40_bench.err; tail -c 120 gpurun_out/r2_bench_c3_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_reference_arm.json 2>> gpurun_out/r40_bench.err; tail -c 200 gpurun_out/r2_bench_c3_reference_arm.json
python tools/profile_step.py --kineto gpurun_out/kernels_c3.txt > /dev/null 2>&1
for w in c1_poc_synthetic c2_cartpole_synthetic c4_minigrid_gtrxl_synthetic c5_mortar_synthetic; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_bench_$w.json 2>> gpurun_out/r40_bench.err; tail -c 100 gpurun_out/r2_bench_$w.json; echo
done
