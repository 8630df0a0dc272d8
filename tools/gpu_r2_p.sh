#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2p_pytest.log
timeout 900 python tools/prof_obb.py 100e6 > gpurun_out/r2p_obb.log 2>&1; tail -5 gpurun_out/r2p_obb.log
timeout 600 python bench.py --workload voxel_geoid --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_voxel_geoid.json 2> gpurun_out/r2p_voxel_geoid.err; echo "voxel_geoid rc=$?"
PCH_GEO_NO_TENSORMAP=1 timeout 600 python bench.py --workload voxel_geoid --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_voxel_geoid_1d.json 2> /dev/null
python - <<'PY'
import json
for f in ('r2p_voxel_geoid','r2p_voxel_geoid_1d'):
    p=json.load(open(f'gpurun_out/{f}.json'))
    print(f, p['value']/1e9, p['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in list(p['kernels'].items())[:5]})
PY
