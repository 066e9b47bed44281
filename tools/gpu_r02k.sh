# round 2, GPU session k: fused map encoder with interleaved points; sanitizers; full bench line
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "map_ or forward_matches or rollout_matches") > gpurun_out/r02k_pytest.log 2>&1; tail -6 gpurun_out/r02k_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02k_bench64.json 2> gpurun_out/r02k_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02k_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved']); print(d['encoder_attn']); print(d['encoder_attn_fused']); print(d['kernel_shares'])"; tail -3 gpurun_out/r02k_bench64.err
bash tools/sanitize.sh r02k
