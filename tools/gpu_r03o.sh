# round 2, last GPU call: HEAD after the host-parser change - smoke + closed-loop fixtures + pipelined evaluation
set -x
mkdir -p gpurun_out
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r03o_smoke.log 2>&1; tail -1 gpurun_out/r03o_smoke.log
(timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout_matches_reference or config2_scene or dt_rollout or pipelined or from_the_reference_file") > gpurun_out/r03o_pytest.log 2>&1; tail -3 gpurun_out/r03o_pytest.log
