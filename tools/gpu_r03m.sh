set -x
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipelined or full_size") > gpurun_out/r03m_pytest.log 2>&1; tail -4 gpurun_out/r03m_pytest.log
