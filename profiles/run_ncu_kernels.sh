#!/bin/bash
# Fast ncu recipe (ONE GPU, ~2 min): `--set full` captures of the hot kernels launched directly by the micro-benchmarks
# at bench-like sizes (no 33-step run-up under the profiler, which is what makes run_ncu.sh take > 10 min).
#   usage: profiles/run_ncu_kernels.sh <tag>  ->  gpurun_out/<tag>_{attn,gemm,small}.ncu-rep
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:attn_tc_kernel -s 2 -c 1 -o gpurun_out/${TAG}_attn -f python tools/attn_bench.py 90 > gpurun_out/${TAG}_attn.log 2>&1
GEMM_MODEL=1 timeout 300 $NCU -k regex:gemm_tc_ta -s 2 -c 1 -o gpurun_out/${TAG}_gemm768 -f python tools/gemm_bench.py 1 > gpurun_out/${TAG}_gemm.log 2>&1
timeout 200 $NCU -k "regex:attn_step_kernel|map_encode_pool_kernel" -s 2 -c 1 -o gpurun_out/${TAG}_attn_step -f python tools/small_kernels_bench.py 90 > gpurun_out/${TAG}_small.log 2>&1
timeout 200 $NCU -k regex:map_encode_pool_kernel -s 2 -c 1 -o gpurun_out/${TAG}_map_encode_pool -f python tools/small_kernels_bench.py 90 >> gpurun_out/${TAG}_small.log 2>&1
ls -la gpurun_out | grep ${TAG}
