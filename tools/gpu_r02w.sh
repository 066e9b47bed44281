# round 2, GPU session w (2 GPUs): the driver's multi-GPU launch of both arms
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --scenes 64 > gpurun_out/r02w_bench2.json 2> gpurun_out/r02w_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r02w_bench2.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['phases'], d['clocks'])"; tail -3 gpurun_out/r02w_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02w_ref2.json 2> gpurun_out/r02w_ref2.err; tail -c 600 gpurun_out/r02w_ref2.json; tail -3 gpurun_out/r02w_ref2.err
