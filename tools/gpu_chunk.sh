#!/bin/bash
# Time-chunked schedule check: whole-network parity for several chunk sizes, the full GPU suite, one bench line.
tag=${1:-chunk}
mkdir -p gpurun_out
for tc in 3 2; do
  echo "== REFID_TCHUNK=$tc"
  REFID_TCHUNK=$tc timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -n 12
done > gpurun_out/pytest_chunk_$tag.log 2>&1
grep -E "^==|passed|failed|Error|error" gpurun_out/pytest_chunk_$tag.log | head -40
timeout 1800 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -n 12 gpurun_out/pytest_gpu_$tag.log
bash tools/gpu_ab.sh chunk_$tag REFID_TCHUNK=0 REFID_TCHUNK=8
