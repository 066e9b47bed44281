set -x
./tools/micro/tmem_bw > gpurun_out/s3_tmem_bw.log 2>&1; cat gpurun_out/s3_tmem_bw.log
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s3_pytest.log 2>&1; tail -5 gpurun_out/s3_pytest.log
python tools/gemm_bench.py > gpurun_out/s3_gemm_bench.log 2>&1; cat gpurun_out/s3_gemm_bench.log
python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s3_bench64.log 2>&1; tail -1 gpurun_out/s3_bench64.log
CTRLSIM_WLO=0 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s3_bench64_nowlo.log 2>&1; tail -1 gpurun_out/s3_bench64_nowlo.log
CTRLSIM_MAP_CACHE=0 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/s3_bench64_nomc.log 2>&1; tail -1 gpurun_out/s3_bench64_nomc.log
