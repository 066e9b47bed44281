set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "linear_res_ln or forward_matches") > gpurun_out/r03i_pytest.log 2>&1; tail -25 gpurun_out/r03i_pytest.log
