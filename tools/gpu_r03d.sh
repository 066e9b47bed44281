set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dt_ or adapter or tracked") > gpurun_out/r03d_pytest.log 2>&1; tail -12 gpurun_out/r03d_pytest.log
