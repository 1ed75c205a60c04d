#!/bin/bash
# N-GPU bench line of the final build (launched the way the driver does).  usage: tools/gpu_r2_scale.sh N
n=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_n${n}_r02.json 2> gpurun_out/bench_n${n}_r02.err; tail -c 1200 gpurun_out/bench_n${n}_r02.json | head -c 600; tail -n 3 gpurun_out/bench_n${n}_r02.err
