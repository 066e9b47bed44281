# round 2: GEMM + residual + LayerNorm fused (gemm_tc_ta_ln_kernel) - closed-loop fixtures and the 64-scene bench, A/B
set -x
mkdir -p gpurun_out
K="rollout_matches_reference or config2_scene or dt_rollout or dt_as_shipped or planner_adversary or caches"
(timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K") > gpurun_out/r03j_pytest.log 2>&1; tail -5 gpurun_out/r03j_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r03j_bench64.json 2> gpurun_out/r03j_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r03j_bench64.json')); print(d['value'], d['phases']['full_window_ms_per_step'], d['phases']['cached_ms_per_step'], d['roofline']['achieved'], d['kernel_shares']['gemm'])"; tail -3 gpurun_out/r03j_bench64.err
CTRLSIM_LNFUSE=0 timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu --no-e2e > gpurun_out/r03j_bench64_nofuse.json 2> gpurun_out/r03j_bench64_nofuse.err; python -c "
import json; d=json.load(open('gpurun_out/r03j_bench64_nofuse.json')); print(d['value'], d['phases']['full_window_ms_per_step'], d['phases']['cached_ms_per_step'], d['roofline']['achieved'], d['kernel_shares']['gemm'])"
