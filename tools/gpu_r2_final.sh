#!/bin/bash
# Round-2 final evidence (one GPU): tests, bench lines of every configuration, layer table, launch probe, ncu launch list and
# `--set full` captures of the chunk-sized EGACA / 1x1 kernels (HBM GB/s).
tag=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 14 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 700 gpurun_out/bench_$tag.json; tail -n 3 gpurun_out/bench_$tag.err
REFID_TCHUNK=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_stepmajor_$tag.json 2>/dev/null; tail -c 300 gpurun_out/bench_stepmajor_$tag.json
timeout 600 python bench.py --workload gopro_11p1_b1 --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_b1_$tag.json 2>&1; tail -c 400 gpurun_out/bench_b1_$tag.json
timeout 600 python bench.py --workload gopro_7skip --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_7skip_$tag.json 2>&1; tail -c 400 gpurun_out/bench_7skip_$tag.json
timeout 600 python bench.py --workload highrev_11p3 --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_highrev_$tag.json 2>&1; tail -c 400 gpurun_out/bench_highrev_$tag.json
timeout 600 python bench.py --workload fullres_720p --steps 5 --warmup 3 > gpurun_out/bench_720p_$tag.json 2> gpurun_out/bench_720p_$tag.err; tail -c 400 gpurun_out/bench_720p_$tag.json
timeout 600 python tools/layer_profile.py 8 23 256 256 layersT23_$tag > gpurun_out/layersT23_$tag.log 2>&1; head -n 3 gpurun_out/layersT23_$tag.log
timeout 300 python tools/host_launch_probe.py > gpurun_out/host_launch_$tag.txt 2>&1; tail -n 1 gpurun_out/host_launch_$tag.txt
timeout 300 python tools/host_launch_probe.py 1 23 256 256 >> gpurun_out/host_launch_$tag.txt 2>&1; tail -n 1 gpurun_out/host_launch_$tag.txt
export REFID_GRAPHS=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_list_$tag.log 2>&1
STEP="python tools/profile_step.py 8 8 256 256"
cap() { # name regex count skip
  timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"$2" -s $4 -c $3 -f -o gpurun_out/prof_$1_$tag $STEP > gpurun_out/ncu_$1_$tag.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1_$tag.ncu-rep gpurun_out/ncu_$1_$tag > /dev/null 2>&1
  rm -f gpurun_out/prof_$1_$tag.ncu-rep
  head -n 16 gpurun_out/ncu_$1_$tag.txt | cut -c1-200
}
cap chunk_elementwise 'k_(ln_fwd|ln_bwd|dw_fwd|dw_bwd|se_fwd|repeat|sum_series|addmask)' 12 6
cap chunk_1x1 'haloconv_kernel<\(int\)(64|128|256), \(int\)[12], \(int\)1' 10 4
du -sh gpurun_out
