# round 2, GPU session v: the driver's own bench invocations (reference arm first, then the product arm), ncu --set full of the
# GEMM launch profiles/ncu_traffic.json describes
set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; tail -c 1500 gpurun_out/r02v_bench_ref.json; tail -4 gpurun_out/r02v_bench_ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02v_bench.json')); print(d['value'], d['e2e']['value'], d['phases'], d['roofline']['achieved'], d['clocks']); print(d['kernel_shares']); print(d.get('cpu_baseline',{}).get('value'), d.get('gpu_torch_baseline',{}).get('value'))"; tail -4 gpurun_out/r02v_bench.err
GEMM_MODEL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ta -s 2 -c 1 -o gpurun_out/r02v_gemm_ffn2 -f python tools/gemm_bench.py ffn2 > gpurun_out/r02v_ncu_gemm.log 2>&1; tail -3 gpurun_out/r02v_ncu_gemm.log
ls -la gpurun_out | grep r02v
