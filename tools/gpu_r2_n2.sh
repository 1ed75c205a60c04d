#!/bin/bash
# Two-GPU evidence: the 2-rank NCCL parity test and the N=2 bench line (both arms' launch contract).
tag=${1:-r02_n2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/pytest_multi_$tag.log 2>&1; tail -n 6 gpurun_out/pytest_multi_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; tail -c 900 gpurun_out/bench_n2_$tag.json; tail -n 3 gpurun_out/bench_n2_$tag.err
