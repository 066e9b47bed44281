# round 2, GPU session x: compute-sanitizer memcheck over a rollout with the TMEM-A GEMM as the default
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py 35 > gpurun_out/r02x_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02x_memcheck.log
tail -5 gpurun_out/r02x_memcheck.log
