#!/bin/bash
# final evidence, part 2: ncu launch list + full captures of the hot kernels, c5 device-resident bench
cd /root/repo
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_step.py --range --rollout-steps 2 --minibatches 2 > gpurun_out/ncu41a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tma_gemm -s 4 -c 2 -o gpurun_out/attn_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu41b.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tma_gemm -c 4 -o gpurun_out/tmagemm_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu41c.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:rollout_fused|tc_conv_gather" -c 4 -o gpurun_out/rollout_r2 -f python tools/profile_step.py --range --rollout-steps 1 --minibatches 0 > gpurun_out/ncu41d.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tc_conv -c 11 -o gpurun_out/tcconv_r2 -f python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu41e.log 2>&1
tail -1 gpurun_out/ncu41?.log; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_c3.csv
timeout 600 python bench.py --workload c5_mortar_synthetic --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_mortar_synthetic.json 2> gpurun_out/r41_bench.err; tail -c 300 gpurun_out/r2_bench_c5_mortar_synthetic.json
