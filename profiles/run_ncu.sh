#!/bin/bash
# ncu recipe (run under gpurun, ONE GPU). Small batch (8 scenes = 90 focal groups), full 32-step windows: bench.py
# --steps 1 times one full-window step (t = 33) after an untimed run-up, so the LAST step of the launch list is the
# steady-state (sliding-window) one.
#   usage: profiles/run_ncu.sh <tag> [extra bench args]
#   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{gemm,attn,attn_step,map_encode_pool}.ncu-rep
TAG=${1:-r02}
shift
set -x
mkdir -p gpurun_out
BENCH="python bench.py --scenes 8 --warmup 1 --steps 1 --no-cpu --no-e2e --no-torch-gpu --chunk 128 $*"
# 1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# 2) full captures of the hot kernel classes: the launches of the last (full-window) step, counted from the launch list
cnt() { grep -c "$1" gpurun_out/${TAG}_launches.csv; }
NG=$(cnt "gemm_tc"); NA=$(cnt "attn_tc_kernel"); NP=$(cnt "map_encode_pool_kernel"); NS=$(cnt "attn_step")
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $((NG-40)) -c 14 -o gpurun_out/${TAG}_gemm -f $BENCH > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s $((NA-8)) -c 8 -o gpurun_out/${TAG}_attn -f $BENCH > gpurun_out/${TAG}_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_encode_pool -s $((NP-1)) -c 1 -o gpurun_out/${TAG}_map_encode_pool -f $BENCH > gpurun_out/${TAG}_map_pool.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_step -s $((NS-5)) -c 5 -o gpurun_out/${TAG}_attn_step -f $BENCH > gpurun_out/${TAG}_attn_step.log 2>&1
ls -la gpurun_out | grep ${TAG}
