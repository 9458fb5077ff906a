#!/bin/bash
# A/B of an environment variable inside ONE gpurun call (same GPU): tools/ab_env.sh VAR "v1 v2 ..." [rounds]
for r in $(seq 1 ${3:-2}); do
  for v in $2; do
    env $1=$v python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        top = {l['layer'].split('grid=')[1]: l['ms_per_launch'] for l in d['roofline']['by_layer'][:5]}
        print('$1=$v round $r: %.0f frames/s  %.2f ms/step  %s' % (d['value'], d['ms_per_step'], top))
"
  done
done
