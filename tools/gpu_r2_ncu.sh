#!/bin/bash
tag=${1:-r02}
# Round-2 ncu session (one GPU): launch list of a training step + `--set full` captures of the kernels VERDICT r1 asked
# for: the dominant halo conv (all 3x3 shapes), both weight-gradient kernels, the tap-GEMM (heads, per-sample conv3), the
# GELU 1x1, the stride-2 `down` data-gradient, and the EGACA memory-bound kernels.  Graph replay off (plain launches).

export REFID_GRAPHS=0
mkdir -p gpurun_out
STEP="python tools/profile_step.py 8 2 256 256"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_list_$tag.log 2>&1
cap() { # name regex count skip
  timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $4 -c $3 -f -o gpurun_out/prof_$1_$tag $STEP > gpurun_out/ncu_$1_$tag.log 2>&1
  # summarise on the box (gpurun copies back at most 64 MiB): small table + the raw metric page, then drop the report
  python tools/ncu_summary.py gpurun_out/prof_$1_$tag.ncu-rep gpurun_out/ncu_$1_$tag > /dev/null 2>&1
  ncu -i gpurun_out/prof_$1_$tag.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/ncu_$1_${tag}_raw.csv.gz
  rm -f gpurun_out/prof_$1_$tag.ncu-rep
  head -n 14 gpurun_out/ncu_$1_$tag.txt | cut -c1-220
}
cap haloconv3x3 'haloconv_kernel<\(int\)(64|128|256), \(int\)[12], \(int\)9, \(int\)64' 8 12
cap haloconv32 'haloconv_kernel<\(int\)(32|64|128), \(int\)[12], \(int\)9, \(int\)32' 4 6
cap haloconv1x1 'haloconv_kernel<\(int\)(64|128|256), \(int\)[12], \(int\)1' 8 10
cap halowgrad 'halowgrad_kernel' 5 2
cap wgrad 'wgrad_kernel' 5 3
cap tapgemm 'tapgemm_kernel' 6 1
cap elementwise 'k_(ln|dw|se|addmask|colsum|sum_series|unroll5|gout_pack|gate_wgrad)' 14 8
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
