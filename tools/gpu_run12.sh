set -x
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_tma_kernel -s 1 -c 1 -o gpurun_out/r01c_gemm768 -f python tools/gemm_trace.py 768 256 > gpurun_out/s12_ncu_gemm.log 2>&1; tail -3 gpurun_out/s12_ncu_gemm.log
