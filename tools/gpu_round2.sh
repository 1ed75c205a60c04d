#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_full.log
REFID_NO_HALO=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_nohalo.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_v2.csv python tools/profile_step.py 2 2 256 256 > gpurun_out/ncu_list.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench_full.log; tail -n 2 gpurun_out/bench_full_nohalo.log
