set -x
./tools/micro/tmem_bw > gpurun_out/s2_tmem_bw.log 2>&1; cat gpurun_out/s2_tmem_bw.log
python tools/attn_err.py > gpurun_out/s2_attn_err.log 2>&1; cat gpurun_out/s2_attn_err.log
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s2_pytest2.log 2>&1; tail -5 gpurun_out/s2_pytest2.log
python tools/attn_bench.py 64 > gpurun_out/s2_attn_bench.log 2>&1; cat gpurun_out/s2_attn_bench.log
python tools/attn_trace.py 64 > gpurun_out/s2_attn_trace.log 2>&1; tail -12 gpurun_out/s2_attn_trace.log
python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s2_bench64.log 2>&1; tail -1 gpurun_out/s2_bench64.log
BENCH="python bench.py --scenes 8 --warmup 34 --steps 2 --no-cpu --no-e2e --chunk 128"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_tma_kernel -s 2570 -c 12 -o gpurun_out/r01b_gemm -f $BENCH > gpurun_out/r01b_gemm.log 2>&1
