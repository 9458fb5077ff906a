# A/B of the once-per-clip branch on its own stream (default) vs serialised (NM_ST_OVERLAP=0): step time and the branch's layers
for m in 1 0; do echo "NM_ST_OVERLAP=$m"; NM_ST_OVERLAP=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'])
for l in d['roofline']['by_layer']:
    if 'n=64 ' in l['layer']: print(' ', l['layer'], l['ms_per_launch'], l['tflops'])"; done
