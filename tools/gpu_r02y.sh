# round 2, GPU session y: whole parity suite + 64-scene bench after the tensor-map cache
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02y_pytest.log 2>&1; tail -4 gpurun_out/r02y_pytest.log
timeout 600 python bench.py --scenes 64 --steps 20 --warmup 3 --no-cpu --no-torch-gpu > gpurun_out/r02y_bench64.json 2> gpurun_out/r02y_bench64.err; python -c "
import json; d=json.load(open('gpurun_out/r02y_bench64.json')); print(d['value'], d['e2e']['value'], d['phases'], d['roofline']['achieved'], d['clocks'])"; tail -3 gpurun_out/r02y_bench64.err
