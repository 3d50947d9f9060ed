cd $GRAFT_REPO_ROOT
for dbg in 0 1 2 4 3 5 6 7; do for bn in 32 128; do echo "== debug $dbg BN $bn"; TRXL_TC_DEBUG=$dbg TRXL_TC_BN=$bn timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "M=300 |M=2048 N=256 K=256|N=256 K=3136"; done; done > gpurun_out/r2_gemm_dbg.log 2>&1
cat gpurun_out/r2_gemm_dbg.log
