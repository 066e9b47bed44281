# round 2, GPU session s: whole parity suite + 64-scene bench with gemm_tc_ta_kernel (A operand in TMEM) as the default GEMM
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02s_pytest.log 2>&1; tail -6 gpurun_out/r02s_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02s_bench64.json 2> gpurun_out/r02s_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02s_bench64.json')); print(d['value'], d['phases'], d['roofline']['achieved']); print(d['kernel_shares'])"; tail -3 gpurun_out/r02s_bench64.err
CTRLSIM_GEMM=tma timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r02s_bench64_tma.json 2> gpurun_out/r02s_bench64_tma.err; python -c "
import json; d=json.load(open('gpurun_out/r02s_bench64_tma.json')); print(d['value'], d['phases'], d['roofline']['achieved']); print(d['kernel_shares'])"; tail -3 gpurun_out/r02s_bench64_tma.err
