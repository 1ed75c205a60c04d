#!/bin/bash
# Parity + per-layer profile in one GPU session.  Usage: bash tools/gpu_check.sh <tag> [bench]
tag=${1:-chk}
mkdir -p gpurun_out
timeout 600 python tools/kernel_probe.py > gpurun_out/kernel_probe_$tag.log 2>&1
tail -n 1 gpurun_out/kernel_probe_$tag.log
grep -v "'ok': True" gpurun_out/kernel_probe_$tag.log | head -20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1
tail -n 3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python tools/layer_profile.py 8 4 256 256 layers_$tag > gpurun_out/layers_$tag.log 2>&1
head -n 1 gpurun_out/layers_$tag.log
if [ "$2" = "bench" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    print("BENCH", round(d["value"], 1), "frames/s", round(d["ms_per_step"], 2), "ms/step", "roofline frac", round(d["roofline"]["frac"], 3), {k: (round(v["ms"], 1), round(v["tflops"])) for k, v in d["roofline"]["per_class"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$tag.log").read()[-2000:])
PY
fi
