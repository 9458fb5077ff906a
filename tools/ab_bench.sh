#!/bin/bash
# A/B of NM_SLAB_DEBUG variants inside ONE gpurun call (same GPU): usage tools/ab_bench.sh "0 32 64" [rounds]
for r in $(seq 1 ${2:-2}); do
  for v in $1; do
    NM_SLAB_DEBUG=$v python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        top = {l['layer'].split('grid=')[1]: l['ms_per_launch'] for l in d['roofline']['by_layer'][:4]}
        print('variant $v round $r: %.0f frames/s  %.2f ms/step  %s' % (d['value'], d['ms_per_step'], top))
"
  done
done
