cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log; cat gpurun_out/r2_tests2.log
for cfg in "1 0" "2 0" "4 0"; do
  set -- $cfg
  echo "=== groups=$1 spin=$2" >> gpurun_out/r2_probe4.log
  TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=$1 TRXL_SPIN_STEPPING=$2 timeout 300 python tools/e2e_probe.py --rollouts 4 >> gpurun_out/r2_probe4.log 2>&1
done
echo "=== groups=2 profile" >> gpurun_out/r2_probe4.log
TRXL_ROLLOUT_GROUPS=2 TRXL_SPIN_STEPPING=0 timeout 300 python tools/e2e_probe.py --rollouts 3 --profile 2>&1 | grep -v Warn | cut -c1-70,150-250 >> gpurun_out/r2_probe4.log
cat gpurun_out/r2_probe4.log
