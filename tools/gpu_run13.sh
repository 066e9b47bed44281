set -x
timeout 300 python tools/gemm_bench.py > gpurun_out/s13_gemm_bench.log 2>&1; cat gpurun_out/s13_gemm_bench.log
timeout 300 python tools/gemm_trace.py 768 256 > gpurun_out/s13_gemm_trace_768.log 2>&1; tail -11 gpurun_out/s13_gemm_trace_768.log
GEMM_DEBUG=2 timeout 300 python tools/gemm_trace.py 768 256 > gpurun_out/s13_gemm_trace_768_nostore.log 2>&1; tail -11 gpurun_out/s13_gemm_trace_768_nostore.log
timeout 120 python tools/attn_bench.py 64 > gpurun_out/s13_attn_bench.log 2>&1; cat gpurun_out/s13_attn_bench.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s13_pytest.log 2>&1; tail -4 gpurun_out/s13_pytest.log
timeout 600 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s13_bench64.log 2>&1; tail -c 700 gpurun_out/s13_bench64.log
