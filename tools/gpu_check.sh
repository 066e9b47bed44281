# GPU check used during development (run under gpurun): parity tests + a short bench line
set -x
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/check_pytest.log 2>&1; tail -6 gpurun_out/check_pytest.log
timeout 600 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/check_bench64.log 2>&1; tail -c 600 gpurun_out/check_bench64.log
