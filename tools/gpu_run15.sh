set -x
timeout 300 python tools/attn_err.py > gpurun_out/s16_attn_err.log 2>&1; cat gpurun_out/s16_attn_err.log
timeout 120 python tools/attn_bench.py 64 > gpurun_out/s16_attn_bench.log 2>&1; cat gpurun_out/s16_attn_bench.log
timeout 120 python tools/attn_trace.py 64 > gpurun_out/s16_attn_trace.log 2>&1; tail -6 gpurun_out/s16_attn_trace.log
