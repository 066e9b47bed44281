# round 2, GPU session o: GEMM with L2 prefetch of the next A tile (A/B by debug bit 64), streaming C stores (bit 128)
set -x
mkdir -p gpurun_out
for d in 0 64 128 192; do GEMM_DEBUG=$d GEMM_MODEL=1 timeout 200 python tools/gemm_bench.py 2>&1 | grep "M=" ; done > gpurun_out/r02o_gemm_prefetch.txt 2>&1
cat gpurun_out/r02o_gemm_prefetch.txt
