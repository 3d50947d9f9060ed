cd $GRAFT_REPO_ROOT
timeout 300 python tools/gemm_bench.py > gpurun_out/r2_gemm_tc.log 2>&1; cat gpurun_out/r2_gemm_tc.log
TRXL_TCGEN05=0 timeout 300 python tools/gemm_bench.py > gpurun_out/r2_gemm_simt.log 2>&1; grep -v "max err" gpurun_out/r2_gemm_simt.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests4.log; cat gpurun_out/r2_tests4.log
