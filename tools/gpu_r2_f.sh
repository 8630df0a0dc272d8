#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/obb_debug.py > gpurun_out/r2f_obb.log 2>&1; cat gpurun_out/r2f_obb.log | tail -30
