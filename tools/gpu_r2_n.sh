#!/bin/bash
mkdir -p gpurun_out
for mc in 2048 8192; do
  PCH_OBB_MAX_CAND=$mc timeout 600 python tools/prof_obb.py 100e6 > gpurun_out/r2n_obb_$mc.log 2>&1; echo "max_cand=$mc"; tail -5 gpurun_out/r2n_obb_$mc.log
done
timeout 600 python -m pytest tests/test_gpu_obb.py -x -q 2>&1 | tail -3
