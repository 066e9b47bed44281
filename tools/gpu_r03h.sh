# round 2: how sensitive are the fixture comparisons to rounding-level changes of the kernels?  The same closed-loop
# fixture tests with the attention on mma.sync (other accumulation order) and with the FP32 FFMA GEMM (other rounding)
set -x
mkdir -p gpurun_out
K="rollout_matches_reference or config2_scene or dt_rollout or dt_as_shipped or planner_adversary or adapter_matches"
(CTRLSIM_ATTN=mma timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K") > gpurun_out/r03h_attn_mma.log 2>&1; tail -15 gpurun_out/r03h_attn_mma.log
(CTRLSIM_GEMM=simt timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$K") > gpurun_out/r03h_gemm_simt.log 2>&1; tail -15 gpurun_out/r03h_gemm_simt.log
