#!/bin/bash
# Last refresh of the headline evidence with the final build: default bench line, step-major line, layer table, launch list.
tag=${1:-r02_last}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 200 gpurun_out/bench_$tag.json
REFID_TCHUNK=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_stepmajor_$tag.json 2>/dev/null
timeout 600 python bench.py --workload gopro_11p1_b1 --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_b1_$tag.json 2>/dev/null
timeout 600 python bench.py --workload highrev_11p3 --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_highrev_$tag.json 2>/dev/null
timeout 600 python bench.py --workload gopro_7skip --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_7skip_$tag.json 2>/dev/null
timeout 600 python bench.py --workload fullres_720p --steps 5 --warmup 3 > gpurun_out/bench_720p_$tag.json 2>/dev/null
timeout 600 python tools/layer_profile.py 8 23 256 256 layersT23_$tag > gpurun_out/layersT23_$tag.log 2>&1; head -n 2 gpurun_out/layersT23_$tag.log
export REFID_GRAPHS=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_step.py 2 4 256 256 > gpurun_out/ncu_list_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv
