# round 2, GPU session t: adapter in DT mode; bench with prefix-cache chunk statistics
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adapter or dt_ or tracked") > gpurun_out/r02t_pytest.log 2>&1; tail -8 gpurun_out/r02t_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02t_bench64.json 2> gpurun_out/r02t_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02t_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved']); print(d['kernel_shares'])"; tail -3 gpurun_out/r02t_bench64.err
