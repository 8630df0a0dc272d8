#!/bin/bash
mkdir -p gpurun_out
PCH_TRACE_TILES=1 timeout 900 python bench.py --workload corridor400M --points 200e6 --tiles 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2j_trace.out 2> gpurun_out/r2j_trace.err
grep "pch tiles" gpurun_out/r2j_trace.out | tail -4
timeout 900 python tools/prof_obb.py 50e6 > gpurun_out/r2j_obb.log 2>&1; tail -8 gpurun_out/r2j_obb.log
bash tools/gpu_scale.sh 1 corridor400M
