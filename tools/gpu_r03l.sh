# round 2, final validation after the last product-code change: whole GPU suite (both builds) + smoke() + driver bench
set -x
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/r03l_pytest.log 2>&1; tail -4 gpurun_out/r03l_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r03l_smoke.log 2>&1; tail -2 gpurun_out/r03l_smoke.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r03l_bench.json 2> gpurun_out/r03l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r03l_bench.json')); print(d['value'], d['e2e']['value'], d['phases'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks']); print(d.get('cpu_baseline',{}).get('value'), d.get('gpu_torch_baseline',{}).get('value'), d['encoder_attn'].get('frac'))"; tail -4 gpurun_out/r03l_bench.err
