# round 2, GPU session d: full parity suite (ABI 5, new attn_step / map_pool / masks) + the wide build
set -x
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -q -rxX) > gpurun_out/r02d_pytest.log 2>&1; tail -25 gpurun_out/r02d_pytest.log
timeout 600 python bench.py --wide --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02d_bench64_wide.json 2> gpurun_out/r02d_bench64_wide.err; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench64_wide.json')); print(d['value'], d['phases'], d['config']['focal_groups_per_step_avg'], d['kernel_shares'])"; tail -3 gpurun_out/r02d_bench64_wide.err
