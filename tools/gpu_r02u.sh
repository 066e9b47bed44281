# round 2, GPU session u: launch list of the cached phase (steps 0..13, 22 scenes = one 256-group chunk)
set -x
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_cached_launches.csv python tools/cached_steps_run.py 22 14 > gpurun_out/r02u_run.log 2>&1; tail -3 gpurun_out/r02u_run.log
python tools/ncu_shares.py gpurun_out/r02u_cached_launches.csv 2 > gpurun_out/r02u_cached_shares.md; cat gpurun_out/r02u_cached_shares.md
