#!/bin/bash
# Round-2 check: parity tests, launch-overhead probe, default bench line.
tag=${1:-r02c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 25 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/host_launch_probe.py > gpurun_out/host_launch_$tag.txt 2>&1; tail -n 2 gpurun_out/host_launch_$tag.txt
timeout 300 python tools/host_launch_probe.py 1 23 256 256 >> gpurun_out/host_launch_$tag.txt 2>&1; tail -n 2 gpurun_out/host_launch_$tag.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 2500 gpurun_out/bench_$tag.json; tail -n 5 gpurun_out/bench_$tag.err
