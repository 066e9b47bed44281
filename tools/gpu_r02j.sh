# round 2, GPU session j: fused map encoder
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "map_ or forward_matches or rollout_matches or caches or config2 or wide_rollout") > gpurun_out/r02j_pytest.log 2>&1; tail -15 gpurun_out/r02j_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02j_bench64.json 2> gpurun_out/r02j_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02j_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved'], d['encoder_attn'], d['kernel_shares'])"; tail -3 gpurun_out/r02j_bench64.err
