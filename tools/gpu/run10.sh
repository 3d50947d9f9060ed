cd $GRAFT_REPO_ROOT
for bn in 32 64 128; do echo "== BN $bn"; TRXL_TC_BN=$bn timeout 300 python tools/gemm_bench.py 2>&1 | grep -v "max err"; done > gpurun_out/r2_gemm_bn.log 2>&1
echo "== SIMT" >> gpurun_out/r2_gemm_bn.log
TRXL_TCGEN05=0 timeout 300 python tools/gemm_bench.py 2>&1 | grep -v "max err" >> gpurun_out/r2_gemm_bn.log
cat gpurun_out/r2_gemm_bn.log
