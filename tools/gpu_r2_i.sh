#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_towers.py tests/test_gpu_obb.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest.log
bash tools/gpu_scale.sh 4 corridor400M corridor1B_geo
