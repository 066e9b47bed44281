# round 2, GPU session f: GEMM feed-path experiments
set -x
mkdir -p gpurun_out
{ for d in 46 44 38; do GEMM_DEBUG=$d timeout 200 python tools/gemm_bench.py 3; done
  for p in 0 1 3; do CTRLSIM_TMA_L2PROMO=$p timeout 200 python tools/gemm_bench.py 3; done; } > gpurun_out/r02f_gemm_experiments.txt 2>&1; cat gpurun_out/r02f_gemm_experiments.txt
