#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.out 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_1gpu.out | tail -1 > gpurun_out/r2_bench_1gpu.json
timeout 600 python bench.py --workload voxel_geoid > gpurun_out/r2_bench_voxel_geoid.out 2>/dev/null; grep '^{' gpurun_out/r2_bench_voxel_geoid.out | tail -1 > gpurun_out/r2_bench_voxel_geoid.json
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_bench_1gpu.json'))
print(p['value']/1e9, p['ms_per_step'], 'e2e', p['e2e']['value']/1e9, p['e2e']['mode'], 'launches', p['gpu_launches'])
print('  modes', {k:(round(v['value']/1e9,3), round(v['ms_per_step'],2), v['stage_info']) for k,v in (p.get('modes') or {}).items()})
print('  roof', p['roofline'])
print('  cpu', p['cpu_baseline'])
print('  clocks', p['clocks'])
p=json.load(open('gpurun_out/r2_bench_voxel_geoid.json')); print('voxel_geoid', p['value']/1e9, p['ms_per_step'])
PY
