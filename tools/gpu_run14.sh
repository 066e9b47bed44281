set -x
BENCH="python bench.py --scenes 8 --warmup 34 --steps 2 --no-cpu --no-e2e --chunk 128"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01c_launches.csv $BENCH > gpurun_out/r01c_launches.log 2>&1
tail -2 gpurun_out/r01c_launches.log | cut -c 1-300
