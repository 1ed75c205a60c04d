#!/bin/bash
# `ncu --set full` of the dominant kernel (3x3 halo conv, 64-channel slabs) with the LEAN epilogue build: tensor-pipe active %.
tag=${1:-r02_lean}
export REFID_GRAPHS=0
mkdir -p gpurun_out
STEP="python tools/profile_step.py 8 2 256 256"
cap() { # name regex count skip
  timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $4 -c $3 -f -o gpurun_out/prof_$1_$tag $STEP > gpurun_out/ncu_$1_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1_$tag.ncu-rep gpurun_out/ncu_$1_$tag > /dev/null 2>&1
  rm -f gpurun_out/prof_$1_$tag.ncu-rep
  head -n 30 gpurun_out/ncu_$1_$tag.txt | cut -c1-200
}
cap haloconv3x3 'haloconv_kernel<\(int\)(64|128|256), \(int\)[12], \(int\)9, \(int\)64' 24 8
