cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,utilization.gpu,pstate --format=csv,noheader -lms 100 > gpurun_out/r2_clk.csv &
SMI=$!
echo "=== groups=2 blocking, profile" > gpurun_out/r2_probe3.log
TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=2 TRXL_SPIN_STEPPING=0 timeout 300 python tools/e2e_probe.py --rollouts 4 --profile >> gpurun_out/r2_probe3.log 2>&1
sleep 1; echo "MARK" >> gpurun_out/r2_clk.csv
echo "=== groups=1 spin (throttled) profile" >> gpurun_out/r2_probe3.log
TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=1 TRXL_SPIN_STEPPING=1 timeout 300 python tools/e2e_probe.py --rollouts 3 --profile >> gpurun_out/r2_probe3.log 2>&1
kill $SMI
cat gpurun_out/r2_probe3.log
sort gpurun_out/r2_clk.csv | uniq -c | sort -rn | head -20
