set -x
cd $GRAFT_REPO_ROOT
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_tests1.log
cat gpurun_out/r2_tests1.log
for cfg in "1 0" "2 0" "2 1" "4 0"; do
  set -- $cfg
  echo "=== groups=$1 spin=$2" >> gpurun_out/r2_probe1.log
  TRXL_E2E_TRACE=1 TRXL_ROLLOUT_GROUPS=$1 TRXL_SPIN_STEPPING=$2 timeout 300 python tools/e2e_probe.py --rollouts 4 >> gpurun_out/r2_probe1.log 2>&1
done
cat gpurun_out/r2_probe1.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
