#!/bin/bash
# ncu recipe (run under gpurun, ONE GPU). Small batch (8 scenes = 90 focal groups), but full 32-step windows: 34 warm-up
# steps bring the episode to t=34 so the profiled launches are the steady-state (sliding-window) ones.
#   usage: profiles/run_ncu.sh <tag>  -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{gemm,attn,map_pool}.ncu-rep
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
BENCH="python bench.py --scenes 8 --warmup 34 --steps 2 --no-cpu --no-e2e --chunk 128"
# 1) every launch of two steady-state steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# 2) full captures of the hot kernel classes (launch-skip counts matching kernels only; steps 0..31 run the short
#    incremental path, so skip counts are taken from the launch list: the last step's launches are captured)
NG=$(grep -c "gemm_tc_tma_kernel" gpurun_out/${TAG}_launches.csv); NA=$(grep -c "attn_tc_kernel" gpurun_out/${TAG}_launches.csv)
NP=$(grep -c "map_pool_kernel" gpurun_out/${TAG}_launches.csv)
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_tma_kernel -s $((NG-40)) -c 12 -o gpurun_out/${TAG}_gemm -f $BENCH > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s $((NA-8)) -c 8 -o gpurun_out/${TAG}_attn -f $BENCH > gpurun_out/${TAG}_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_pool_kernel -s $((NP-1)) -c 1 -o gpurun_out/${TAG}_map_pool -f $BENCH > gpurun_out/${TAG}_map_pool.log 2>&1
ls -la gpurun_out | grep ${TAG}
