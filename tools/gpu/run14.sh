cd $GRAFT_REPO_ROOT
for bn in 32 64 128; do echo "== BN $bn"; TRXL_TC_BN=$bn timeout 300 python tools/gemm_bench.py 2>&1 | head -12; done > gpurun_out/r2_gemm_bn3.log 2>&1
cat gpurun_out/r2_gemm_bn3.log
echo "== debug 7 (floor)"; for bn in 32 128; do TRXL_TC_DEBUG=7 TRXL_TC_BN=$bn timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "M=300 "; done
