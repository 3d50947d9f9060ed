cd $GRAFT_REPO_ROOT
TRXL_TC_BN=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_gemm -s 3 -c 2 -o gpurun_out/r2_tmagemm_bn32 python tools/gemm_one.py 2048 256 256 fwd > gpurun_out/ncu11.log 2>&1
TRXL_TC_BN=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_gemm -s 3 -c 2 -o gpurun_out/r2_tmagemm_bn128 python tools/gemm_one.py 2048 256 256 fwd >> gpurun_out/ncu11.log 2>&1
tail -5 gpurun_out/ncu11.log; ls -la gpurun_out/
