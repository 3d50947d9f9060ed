cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 40 -c 1 -o gpurun_out/r2_rollout_fused python tools/profile_step.py --range --rollout-steps 60 --minibatches 0 > gpurun_out/ncu17.log 2>&1
tail -3 gpurun_out/ncu17.log; ls -la gpurun_out/*.ncu-rep
