#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py -x -q > gpurun_out/r2c_tiles.log 2>&1; echo "tiles rc=$?"; tail -25 gpurun_out/r2c_tiles.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_tiles.py > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c_bench.json'))
print(p['value']/1e9, p['ms_per_step'])
print({k:round(v['ms_per_step'],3) for k,v in list(p['kernels'].items())[:14]})
PY
