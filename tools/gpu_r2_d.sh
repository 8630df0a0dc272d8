#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err
timeout 900 python bench.py --workload corridor400M --points 80e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_tiled_small.json 2> gpurun_out/r2d_tiled_small.err; echo "tiled rc=$?"; tail -5 gpurun_out/r2d_tiled_small.err
timeout 900 python bench.py --workload corridor1B_geo --points 80e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_geo_small.json 2> gpurun_out/r2d_geo_small.err; echo "geo rc=$?"; tail -5 gpurun_out/r2d_geo_small.err
python - <<'PY'
import json
for f in ('r2d_bench','r2d_tiled_small','r2d_geo_small'):
    try:
        p=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, p['value']/1e9, p['ms_per_step'], p.get('stage_info'), (p.get('e2e') or {}).get('value'))
    print('  modes', {k:(round(v['value']/1e9,2), round(v['ms_per_step'],2), v['stage_info']) for k,v in (p.get('modes') or {}).items()})
    print('  ', {k:round(v['ms_per_step'],3) for k,v in list(p['kernels'].items())[:16]})
    print('  roof', p['roofline']['kernel'], round(p['roofline']['frac'],3), p.get('collectives'))
PY
