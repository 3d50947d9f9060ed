cd $GRAFT_REPO_ROOT
# launch list (shares) of a slice: 2 rollout steps + 2 minibatch steps
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_step.py --range --rollout-steps 2 --minibatches 2 > gpurun_out/ncu23a.log 2>&1
# the two grouped attention GEMMs of block 0 (5th and 6th tma_gemm launches of the profiled minibatch step)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tma_gemm -s 4 -c 2 -o gpurun_out/attn_r2 python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu23b.log 2>&1
# trunk GEMMs: first four tma_gemm launches (lin_hidden, embedding, Wq, per-head K fold)
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tma_gemm -c 4 -o gpurun_out/tmagemm_r2 python tools/profile_step.py --range --rollout-steps 0 --minibatches 1 > gpurun_out/ncu23c.log 2>&1
tail -2 gpurun_out/ncu23a.log gpurun_out/ncu23b.log gpurun_out/ncu23c.log; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_c3.csv
