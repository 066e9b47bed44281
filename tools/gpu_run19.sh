set -x
(time python bench.py) > gpurun_out/s19_bench_default.log 2>&1; tail -c 400 gpurun_out/s19_bench_default.log
bash profiles/run_ncu.sh r01d > gpurun_out/s19_ncu.log 2>&1; tail -8 gpurun_out/s19_ncu.log
