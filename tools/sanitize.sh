# compute-sanitizer memcheck over a small rollout (run under gpurun). Exit code 9 = memcheck reported an error.
set -x
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python tools/sanitize_target.py 35 > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log; tail -5 gpurun_out/sanitize_memcheck.log
