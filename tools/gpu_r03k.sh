set -x
mkdir -p gpurun_out
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r03k_bench64.json 2> gpurun_out/r03k_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r03k_bench64.json')); print(d['value'], d['phases']['full_window_ms_per_step'], d['phases']['cached_ms_per_step'], d['roofline']['achieved'], d['kernel_shares']['gemm'])"; tail -3 gpurun_out/r03k_bench64.err
CTRLSIM_LNFUSE=0 timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r03k_bench64_nofuse.json 2> gpurun_out/r03k_bench64_nofuse.err; python -c "
import json; d=json.load(open('gpurun_out/r03k_bench64_nofuse.json')); print(d['value'], d['phases']['full_window_ms_per_step'], d['phases']['cached_ms_per_step'], d['roofline']['achieved'], d['kernel_shares']['gemm'])"
