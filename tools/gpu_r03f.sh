set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipelined") > gpurun_out/r03f_pytest.log 2>&1; tail -5 gpurun_out/r03f_pytest.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-torch-gpu ) > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r03f_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['episode_wall_s'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e']['metrics'])"; tail -4 gpurun_out/r03f_bench.err
