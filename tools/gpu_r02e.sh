# round 2, GPU session e: reference-signature adapter on the GPU; GEMM bottleneck experiments (debug bits, results wrong)
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "adapter") > gpurun_out/r02e_pytest.log 2>&1; tail -8 gpurun_out/r02e_pytest.log
for d in 0 2 4 8 12 14; do GEMM_DEBUG=$d timeout 200 python tools/gemm_bench.py 3; done > gpurun_out/r02e_gemm_experiments.txt 2>&1; cat gpurun_out/r02e_gemm_experiments.txt
