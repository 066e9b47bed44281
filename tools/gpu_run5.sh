set -x
timeout 300 python tools/attn_err.py > gpurun_out/s5_attn_err.log 2>&1; cat gpurun_out/s5_attn_err.log
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s5_pytest.log 2>&1; tail -5 gpurun_out/s5_pytest.log
timeout 120 python tools/attn_bench.py 64 > gpurun_out/s5_attn_bench.log 2>&1; cat gpurun_out/s5_attn_bench.log
timeout 120 python tools/attn_trace.py 64 > gpurun_out/s5_attn_trace.log 2>&1; tail -5 gpurun_out/s5_attn_trace.log
timeout 600 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s5_bench64.log 2>&1; tail -1 gpurun_out/s5_bench64.log
