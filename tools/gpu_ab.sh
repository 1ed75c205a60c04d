#!/bin/bash
# A/B of an engine switch on the full training step: usage tools/gpu_ab.sh TAG "ENV=a" "ENV=b" ...
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 12 --warmup 3 --no-ref-cuda --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f  ms/step %.2f  e2e %.1f  roofline.frac %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], d['clocks']['sm_mhz']))
print('  per_class', {k: (round(v['ms'],2), round(v['tflops'])) for k,v in r['per_class'].items()})
print('  parity', d.get('parity'))
"
done > gpurun_out/ab_$tag.txt 2>&1
cat gpurun_out/ab_$tag.txt
