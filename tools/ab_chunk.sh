#!/bin/bash
# A/B of NM_FRAME_CHUNK values inside ONE gpurun call (same GPU)
for r in 1 2; do
  for v in $1; do
    NM_FRAME_CHUNK=$v python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('chunk $v round $r: %.0f frames/s  %.2f ms/step e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
"
  done
done
