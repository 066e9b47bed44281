# round 2, GPU session h: packed weight blocks fetched by linear bulk copies
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "linear or forward_matches") > gpurun_out/r02h_pytest.log 2>&1; tail -8 gpurun_out/r02h_pytest.log
{ CTRLSIM_GEMM_PACKED=0 GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py; GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py;
  for d in 2 46; do GEMM_DEBUG=$d GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py 3; done; } > gpurun_out/r02h_gemm_packed.txt 2>&1; grep -v "^+" gpurun_out/r02h_gemm_packed.txt
