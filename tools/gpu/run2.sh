cd $GRAFT_REPO_ROOT
timeout 300 python tools/pcie_probe.py > gpurun_out/r2_pcie.log 2>&1; cat gpurun_out/r2_pcie.log
timeout 600 python tools/c3_minibatch_diag.py > gpurun_out/r2_c3diag.log 2>&1; tail -60 gpurun_out/r2_c3diag.log
