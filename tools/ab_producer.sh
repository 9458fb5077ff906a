for m in 0 512; do echo "DEBUG=$m"; NM_SLAB_DEBUG=$m timeout 300 python bench.py --clips 16 --steps 3 --warmup 3 --no-sub --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'])
for l in d['roofline']['by_layer']:
    if ('grid=64' in l['layer'] or '128->64' in l['layer']) and 'k3' in l['layer']: print(l['layer'], l['ms_per_launch'], l['tflops'])"; done
