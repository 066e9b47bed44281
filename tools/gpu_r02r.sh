# round 2, GPU session r: power / clock probe of the hot kernel classes
set -x
mkdir -p gpurun_out
timeout 300 python tools/power_probe.py > gpurun_out/r02r_power_probe.txt 2>&1; grep -v Warn gpurun_out/r02r_power_probe.txt | tail -20
