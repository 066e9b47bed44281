# GPU check of the contact path (run under gpurun): full parity suite with contacts on, then the simulator-facing tests
# with the contact-free fallback (CTRLSIM_CONTACTS=0), then a short bench line
set -x
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/check_pytest.log 2>&1; tail -15 gpurun_out/check_pytest.log
CTRLSIM_CONTACTS=0 timeout 300 python -m pytest tests -m gpu -q -k "sparse or multi_scene or caches or nucleus or batch_composition" > gpurun_out/check_pytest_nocontacts.log 2>&1; tail -3 gpurun_out/check_pytest_nocontacts.log
timeout 300 python bench.py --scenes 64 --steps 45 --warmup 3 --no-cpu --no-e2e > gpurun_out/check_bench64.log 2>&1; tail -c 300 gpurun_out/check_bench64.log
