#!/bin/bash
# Round-1 ncu recipe (run under gpurun, ONE GPU). Small batch (8 scenes), but full 32-step windows: 34 warm-up steps
# bring the episode to t=34 so the profiled launches are the steady-state (sliding-window) ones.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --scenes 8 --warmup 34 --steps 2 --no-cpu --no-e2e --chunk 128"
# 1) every launch of two steady-state steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 420 --csv --log-file gpurun_out/r01_launches.csv $BENCH > gpurun_out/r01_launches.log 2>&1
# 2) full captures of the three hot kernel classes
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 300 -c 2 -o gpurun_out/r01_gemm -f $BENCH > gpurun_out/r01_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_causal_kernel -s 136 -c 1 -o gpurun_out/r01_attn_causal -f $BENCH > gpurun_out/r01_attn_causal.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_pool_kernel -s 34 -c 1 -o gpurun_out/r01_map_pool -f $BENCH > gpurun_out/r01_map_pool.log 2>&1
ls -la gpurun_out
