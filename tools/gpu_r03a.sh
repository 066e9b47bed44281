set -x
mkdir -p gpurun_out
timeout 300 python tools/host_time_probe.py 64 > gpurun_out/r03a_host_probe.txt 2>&1; grep "t=" gpurun_out/r03a_host_probe.txt
