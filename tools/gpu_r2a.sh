#!/bin/bash
# Round-2 first GPU session: parity tests, bench lines (train / infer / reference arms), launch-overhead probe, layer table.
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 25 gpurun_out/pytest_gpu_$tag.log
grep -h "max-abs\|dPSNR\|sampled-gradient\|MULTI_OK" gpurun_out/pytest_gpu_$tag.log | head -20
timeout 300 python tools/host_launch_probe.py > gpurun_out/host_launch_$tag.txt 2>&1; tail -n 2 gpurun_out/host_launch_$tag.txt
timeout 300 python tools/host_launch_probe.py 1 23 256 256 >> gpurun_out/host_launch_$tag.txt 2>&1; tail -n 2 gpurun_out/host_launch_$tag.txt
REFID_GRAPHS=0 timeout 300 python tools/host_launch_probe.py 1 23 256 256 >> gpurun_out/host_launch_$tag.txt 2>&1; tail -n 1 gpurun_out/host_launch_$tag.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1500 gpurun_out/bench_$tag.json; tail -n 5 gpurun_out/bench_$tag.err
timeout 600 python bench.py --workload gopro_11p1_b1 --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_b1_$tag.json 2>&1; tail -c 600 gpurun_out/bench_b1_$tag.json
timeout 600 python bench.py --workload fullres_720p --steps 5 --warmup 3 > gpurun_out/bench_720p_$tag.json 2> gpurun_out/bench_720p_$tag.err; tail -c 900 gpurun_out/bench_720p_$tag.json; tail -n 3 gpurun_out/bench_720p_$tag.err
timeout 600 python bench.py --mode infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_$tag.json 2>&1; tail -c 500 gpurun_out/bench_infer_$tag.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>&1; tail -c 700 gpurun_out/bench_ref_$tag.json
timeout 600 python bench.py --impl reference-cuda --steps 3 --warmup 2 > gpurun_out/bench_refcuda_$tag.json 2>&1; tail -c 900 gpurun_out/bench_refcuda_$tag.json
timeout 600 python tools/layer_profile.py 8 23 256 256 layersT23_$tag > gpurun_out/layersT23_$tag.log 2>&1; head -n 30 gpurun_out/layersT23_$tag.log
