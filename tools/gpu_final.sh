#!/bin/bash
# Final evidence run: tests, full bench (with CPU baseline), reference arm, launch list, full ncu captures, T=23 layer table.
tag=${1:-r01_final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 2 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; tail -c 400 gpurun_out/bench_ref_$tag.json
timeout 600 python tools/layer_profile.py 8 23 256 256 layersT23_$tag > gpurun_out/layersT23_$tag.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_list_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"haloconv_kernel<\(int\)(64|128|256), \(int\)[12], \(int\)9, \(int\)64" -s 10 -c 7 -f -o gpurun_out/prof_haloconv_$tag python tools/profile_step.py 8 2 256 256 > gpurun_out/ncu_halo_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:halowgrad -c 3 -f -o gpurun_out/prof_halowgrad_$tag python tools/profile_step.py 8 2 256 256 > gpurun_out/ncu_wgrad_$tag.log 2>&1
for u in umma_rate umma_pattern umma_smem_contention; do timeout 120 ./tools/ubench/$u > gpurun_out/${u}_$tag.txt 2>&1; done
timeout 300 python tools/host_launch_probe.py > gpurun_out/host_launch_$tag.txt 2>&1; tail -n 1 gpurun_out/host_launch_$tag.txt
timeout 300 python tools/callers_bench.py > gpurun_out/callers_bench_$tag.txt 2>&1
ls -la gpurun_out/*.ncu-rep
