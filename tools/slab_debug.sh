for m in 0 1 2 4 6 7; do echo "DEBUG=$m"; NM_SLAB_DEBUG=$m timeout 300 python bench.py --clips 6 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for l in d['roofline']['by_layer']:
    if 'grid=64' in l['layer'] and 'k3' in l['layer']: print(l['layer'], l['ms_per_launch'])"; done
