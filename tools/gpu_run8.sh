set -x
timeout 300 python tools/gemm_trace.py 768 256 > gpurun_out/s8_gemm_trace_768.log 2>&1; tail -30 gpurun_out/s8_gemm_trace_768.log
timeout 300 python tools/gemm_trace.py 256 1024 > gpurun_out/s8_gemm_trace_k1024.log 2>&1; tail -12 gpurun_out/s8_gemm_trace_k1024.log
TRACE_WLO=0 timeout 300 python tools/gemm_trace.py 768 256 > gpurun_out/s8_gemm_trace_768_nowlo.log 2>&1; tail -5 gpurun_out/s8_gemm_trace_768_nowlo.log
