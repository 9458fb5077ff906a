#!/bin/bash
# A/B of two builds of the library inside ONE gpurun call (same GPU): csrc/libnm_base.so vs csrc/libnm_new.so
cd "$(dirname "$0")/.."
L=neural_marionette_b200/csrc
for r in $(seq 1 ${1:-2}); do
  for v in base new; do
    cp $L/libnm_$v.so $L/libnm_b200.so
    python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        top = {l['layer'].split('grid=')[1]: l['ms_per_launch'] for l in d['roofline']['by_layer'][:4]}
        print('$v round $r: %.0f frames/s  %.2f ms/step  %s' % (d['value'], d['ms_per_step'], top))
"
  done
done
cp $L/libnm_new.so $L/libnm_b200.so
