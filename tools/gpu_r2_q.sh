#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_obb.py tests/test_gpu_towers.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2q_bench.out 2>/dev/null; grep '^{' gpurun_out/r2q_bench.out | tail -1 > gpurun_out/r2q_bench.json
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2q_bench.json'))
print(p['value']/1e9, p['ms_per_step'])
print('  modes', {k:(round(v['value']/1e9,3), round(v['ms_per_step'],2), v['stage_info']) for k,v in (p.get('modes') or {}).items()})
PY
