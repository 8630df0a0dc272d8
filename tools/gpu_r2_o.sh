#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_obb.py tests/test_gpu_towers.py -x -q 2>&1 | tail -4
timeout 600 python tools/obb_debug.py > gpurun_out/r2o_dbg.log 2>&1; grep "^cluster" gpurun_out/r2o_dbg.log | cut -c1-210
timeout 900 python tools/prof_obb.py 100e6 > gpurun_out/r2o_obb.log 2>&1; tail -6 gpurun_out/r2o_obb.log
