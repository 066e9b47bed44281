# Round-end measurement (run under gpurun, one GPU): default bench line + ncu launch list / captures of the final kernels
set -x
(time python bench.py) > gpurun_out/final_bench_default.log 2>&1; tail -c 400 gpurun_out/final_bench_default.log
bash profiles/run_ncu.sh r01f > gpurun_out/final_ncu.log 2>&1; tail -6 gpurun_out/final_ncu.log
