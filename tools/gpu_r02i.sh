# round 2, GPU session i: config-2 fixture, replay table, config 4 record, whole suite
set -x
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -q -rxX) > gpurun_out/r02i_pytest.log 2>&1; tail -25 gpurun_out/r02i_pytest.log
timeout 900 python tools/replay_eval.py --scenes 1000 > gpurun_out/r02i_replay_1000.json 2> gpurun_out/r02i_replay_1000.err; head -c 2500 gpurun_out/r02i_replay_1000.json; tail -3 gpurun_out/r02i_replay_1000.err
