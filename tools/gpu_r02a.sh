# round 2, GPU session a: parity suite under the new defaults (glibc trig), the driver's bench line, ncu baseline
set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x -rxX) > gpurun_out/r02a_pytest.log 2>&1; tail -15 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 1500 bash profiles/run_ncu.sh r02a
