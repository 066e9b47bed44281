set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --scenes 64 > gpurun_out/r03g_bench2.json 2> gpurun_out/r03g_bench2.err; grep '^{' gpurun_out/r03g_bench2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['metrics']['collision_rate'], d['clocks'])"; tail -3 gpurun_out/r03g_bench2.err
