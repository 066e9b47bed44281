# round 2, GPU session n: attention with P V of tile j-1 overlapped with the sweep of tile j
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "attn or forward_matches or rollout_matches or caches") > gpurun_out/r02n_pytest.log 2>&1; tail -6 gpurun_out/r02n_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02n_bench64.json 2> gpurun_out/r02n_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02n_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved']); print(d['kernel_shares'])"; tail -3 gpurun_out/r02n_bench64.err
timeout 120 python tools/attn_trace.py > gpurun_out/r02n_attn_trace.txt 2>&1; tail -8 gpurun_out/r02n_attn_trace.txt
timeout 120 python tools/attn_bench.py > gpurun_out/r02n_attn_bench.txt 2>&1; tail -4 gpurun_out/r02n_attn_bench.txt
