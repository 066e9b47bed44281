# compute-sanitizer over a small rollout (run under gpurun): memcheck, then racecheck (shared-memory hazards).
# Exit code 9 = the tool reported an error.  usage: tools/sanitize.sh <tag>
TAG=${1:-r02}
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python tools/sanitize_target.py 35 > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/${TAG}_memcheck.log; tail -5 gpurun_out/${TAG}_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python tools/sanitize_target.py 33 > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/${TAG}_racecheck.log
tail -8 gpurun_out/${TAG}_racecheck.log
