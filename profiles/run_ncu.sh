#!/bin/bash
# ncu recipe (run under gpurun, ONE GPU). Small batch (8 scenes = 90 focal groups), but full 32-step windows: 34 warm-up
# steps bring the episode to t=34 so the profiled launches are the steady-state (sliding-window) ones.
#   usage: profiles/run_ncu.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{gemm,attn_self,attn_cross,map_pool}.ncu-rep
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
BENCH="python bench.py --scenes 8 --warmup 34 --steps 2 --no-cpu --no-e2e --chunk 128"
# 1) every launch of two steady-state steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# 2) full captures of the hot kernel classes (launch-skip counts matching kernels only)
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_tma_kernel -s 2700 -c 3 -o gpurun_out/${TAG}_gemm -f $BENCH > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 272 -c 8 -o gpurun_out/${TAG}_attn -f $BENCH > gpurun_out/${TAG}_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_pool_kernel -s 34 -c 1 -o gpurun_out/${TAG}_map_pool -f $BENCH > gpurun_out/${TAG}_map_pool.log 2>&1
ls -la gpurun_out
