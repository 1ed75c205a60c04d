#!/bin/bash
# ncu captures: launch list of one training step + full-set captures of the dominant kernels (one GPU, short command).
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_list_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:haloconv -s 150 -c 4 -f -o gpurun_out/prof_haloconv_$tag python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_halo_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 2 -c 3 -f -o gpurun_out/prof_wgrad_$tag python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_wgrad_$tag.log 2>&1
ls -la gpurun_out/*.ncu-rep
