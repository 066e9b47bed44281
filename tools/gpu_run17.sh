set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s17_pytest.log 2>&1; tail -4 gpurun_out/s17_pytest.log
timeout 600 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s17_bench64.log 2>&1; tail -c 600 gpurun_out/s17_bench64.log
