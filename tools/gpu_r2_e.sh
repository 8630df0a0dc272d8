#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_obb.py tests/test_gpu_towers.py tests/test_gpu_dropin.py tests/test_gpu_tiles.py -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2e_bench.json'))
print(p['value']/1e9, p['ms_per_step'])
print('  modes', {k:(round(v['value']/1e9,3), round(v['ms_per_step'],2), v['stage_info']) for k,v in (p.get('modes') or {}).items()})
PY
