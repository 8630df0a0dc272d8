#!/bin/bash
mkdir -p gpurun_out
N=4
PCH_TRACE_TILES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --workload corridor400M --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2k_trace4.out 2> gpurun_out/r2k_trace4.err
grep "pch tiles" gpurun_out/r2k_trace4.out | tail -4
bash tools/gpu_scale.sh 4 pipeline
