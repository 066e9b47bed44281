set -x
nvidia-smi -L
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s2_pytest.log 2>&1; tail -5 gpurun_out/s2_pytest.log
(time python bench.py) > gpurun_out/s2_bench_default.log 2>&1; tail -3 gpurun_out/s2_bench_default.log
bash profiles/run_ncu.sh r01b > gpurun_out/s2_ncu.log 2>&1
