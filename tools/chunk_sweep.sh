for c in 320 640 1280; do echo "CHUNK=$c"; NM_FRAME_CHUNK=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'])"; python - <<'PY'
import torch
PY
done
