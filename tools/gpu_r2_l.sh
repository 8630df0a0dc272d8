#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2l_pytest.log
timeout 900 python tools/prof_obb.py 100e6 > gpurun_out/r2l_obb.log 2>&1; tail -6 gpurun_out/r2l_obb.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2l_bench.json'))
print(p['value']/1e9, p['ms_per_step'], 'e2e', p['e2e']['value']/1e9, p['e2e']['mode'])
print('  modes', {k:(round(v['value']/1e9,3), round(v['ms_per_step'],2), v['stage_info']) for k,v in (p.get('modes') or {}).items()})
print('  ', {k:round(v['ms_per_step'],3) for k,v in list(p['kernels'].items())[:16]})
print('  roof', p['roofline']['kernel'], round(p['roofline']['frac'],3), p['roofline']['launches_per_step'])
PY
