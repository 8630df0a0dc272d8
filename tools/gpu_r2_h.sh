#!/bin/bash
# 2 GPUs: NCCL halo exchange check, corridor400M, corridor1B_geo (reduced) and the default pipeline at N=2
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/nccl_halo_check.py > gpurun_out/r2h_nccl_check_$N.log 2>&1; echo "nccl check rc=$?"; grep "nccl halo" gpurun_out/r2h_nccl_check_$N.log || tail -20 gpurun_out/r2h_nccl_check_$N.log
timeout 1200 $TR bench.py --gpus $N --workload corridor400M --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2h_corridor400M_${N}gpu.json 2> gpurun_out/r2h_corridor400M_${N}gpu.err; echo "400M rc=$?"; tail -3 gpurun_out/r2h_corridor400M_${N}gpu.err
python - <<PY
import json
try:
    p=json.load(open('gpurun_out/r2h_corridor400M_${N}gpu.json'))
    print('400M', p['n_gpus'], p['value']/1e9, p['ms_per_step'], p['stage_info'], p['collectives'], (p.get('e2e') or {}).get('value'))
except Exception as e: print('unreadable', e)
PY
