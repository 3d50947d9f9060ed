cd $GRAFT_REPO_ROOT
for cfg in "1 0" "2 0"; do
  set -- $cfg
  echo "=== groups=$1 spin=$2" >> gpurun_out/r2_probe2.log
  TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=$1 TRXL_SPIN_STEPPING=$2 timeout 300 python tools/e2e_probe.py --rollouts 4 >> gpurun_out/r2_probe2.log 2>&1
done
echo "=== groups=2 spin=0 nographs" >> gpurun_out/r2_probe2.log
TRXL_NO_GRAPHS=1 TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=2 TRXL_SPIN_STEPPING=0 timeout 300 python tools/e2e_probe.py --rollouts 3 >> gpurun_out/r2_probe2.log 2>&1
cat gpurun_out/r2_probe2.log
timeout 900 python tools/c3_minibatch_diag.py > gpurun_out/r2_c3diag2.log 2>&1; grep -v "^transformer.transformer_blocks.[12]" gpurun_out/r2_c3diag2.log | tail -80
