set -x
(time python bench.py) > gpurun_out/s7_bench_default.log 2>&1; tail -1 gpurun_out/s7_bench_default.log
(time python bench.py --impl reference --steps 4 --warmup 1) > gpurun_out/s7_bench_ref.log 2>&1; tail -2 gpurun_out/s7_bench_ref.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
