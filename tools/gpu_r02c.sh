# round 2, GPU session c: attn_step on mma.sync, GEMM with a 4-deep TMA ring
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "linear or attn_step or map_pool or forward_matches or rollout_matches or caches") > gpurun_out/r02c_pytest.log 2>&1; tail -15 gpurun_out/r02c_pytest.log
timeout 300 python tools/small_kernels_bench.py 90 > gpurun_out/r02c_small_kernels.txt 2>&1; cat gpurun_out/r02c_small_kernels.txt
timeout 300 python tools/gemm_bench.py > gpurun_out/r02c_gemm_bench.txt 2>&1; cat gpurun_out/r02c_gemm_bench.txt
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02c_bench64.json 2> gpurun_out/r02c_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved'], d['encoder_attn'].get('frac'), d['kernel_shares'])"; tail -3 gpurun_out/r02c_bench64.err
