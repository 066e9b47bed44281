# round 2, GPU session z: GEMM epilogue refactor check
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "linear or forward_matches or rollout_matches or wide_rollout") > gpurun_out/r02z_pytest.log 2>&1; tail -4 gpurun_out/r02z_pytest.log
(GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py; timeout 200 python tools/gemm_bench.py) 2>&1 | grep "M=" 
