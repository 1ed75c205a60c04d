#!/bin/bash
# One GPU session: kernel parity, halo diagnostics, network parity, tests, small bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for bo in 0 1 2; do
  REFID_HALO_BO=$bo timeout 300 python tools/halo_tap_probe.py > gpurun_out/halo_tap_bo$bo.log 2>&1
done
timeout 600 python tools/kernel_probe.py > gpurun_out/kernel_probe.log 2>&1
REFID_NO_HALO=1 timeout 600 python tools/net_probe.py blurry_t3_32 > gpurun_out/net_probe_nohalo.log 2>&1
timeout 600 python tools/net_probe.py blurry_t3_32 > gpurun_out/net_probe_halo.log 2>&1
tail -3 gpurun_out/halo_tap_bo*.log gpurun_out/kernel_probe.log
grep -E "out err|abort|grad norm rel" gpurun_out/net_probe_*.log
