# round 2, GPU session q: GEMM with the A operand in TMEM (gemm_tc_ta_kernel) vs the smem-A kernel; N1 third mode test
set -x
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "linear or tracked_rtgs") > gpurun_out/r02q_pytest.log 2>&1; tail -15 gpurun_out/r02q_pytest.log
(GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py; CTRLSIM_GEMM=tma GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py; GEMM_DEBUG=2 GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py) > gpurun_out/r02q_gemm.txt 2>&1; cat gpurun_out/r02q_gemm.txt | grep -v Warn | tail -20
