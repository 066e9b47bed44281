# round 2, GPU session p: N1 (decision transformer / real-time rewards) product path vs the reference fixture
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dt_") > gpurun_out/r02p_pytest.log 2>&1; tail -30 gpurun_out/r02p_pytest.log
