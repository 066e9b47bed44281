# First GPU session of the next round (run under gpurun): what could not be run on a B200 after this round's GPU budget
# was spent.  1) the whole parity suite (the in-contact tolerances of tests/parity_checks.py were set from the CPU
# emulation of the GPU arithmetic; the two tests concerned have since passed on a B200); 2) the same suite with glibc's
# trig on the GPU (the crowded episode is already bit-exact there) - expected: everything stays green, then flip the
# default in api.cu;
# 3) the headline bench.
set -x
(time timeout 900 python -m pytest tests -m gpu -q -rxX) > gpurun_out/next_pytest.log 2>&1; tail -8 gpurun_out/next_pytest.log
CTRLSIM_TRIG=glibc timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/next_pytest_glibc_trig.log 2>&1; tail -8 gpurun_out/next_pytest_glibc_trig.log
timeout 600 python bench.py > gpurun_out/next_bench.log 2>&1; tail -c 1500 gpurun_out/next_bench.log
