# round 2, GPU session b: L2 -> SM bandwidth microbenchmark; new attn_step + map_pool kernels (parity + speed)
set -x
mkdir -p gpurun_out
timeout 300 tools/micro/l2_tma_bw > gpurun_out/r02b_l2_tma_bw.txt 2>&1; cat gpurun_out/r02b_l2_tma_bw.txt
(timeout 900 python -m pytest tests -m gpu -q -x -k "attn_step or map_pool or forward_matches or rollout_matches or caches") > gpurun_out/r02b_pytest.log 2>&1; tail -15 gpurun_out/r02b_pytest.log
timeout 300 python tools/small_kernels_bench.py 90 > gpurun_out/r02b_small_kernels.txt 2>&1; cat gpurun_out/r02b_small_kernels.txt
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02b_bench64.json 2> gpurun_out/r02b_bench64.err; tail -c 2500 gpurun_out/r02b_bench64.json; tail -3 gpurun_out/r02b_bench64.err
