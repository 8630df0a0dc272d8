#!/bin/bash
# usage: gpu_scale.sh N workload [workload...]   (one bench line per workload at N GPUs -> gpurun_out/r2_<workload>_<N>gpu.json)
mkdir -p gpurun_out
N=$1; shift
for W in "$@"; do
  EXTRA=""
  case $W in pipeline) EXTRA="--steps 10 --warmup 3 --no-modes";; *) EXTRA="--steps 3 --warmup 2";; esac
  if [ "$N" = 1 ]; then CMD="python bench.py --gpus 1"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N"; fi
  timeout 2400 $CMD --workload $W $EXTRA --no-cpu-baseline > gpurun_out/r2_${W}_${N}gpu.out 2> gpurun_out/r2_${W}_${N}gpu.err; echo "$W @ $N rc=$?"
  grep '^{' gpurun_out/r2_${W}_${N}gpu.out | tail -1 > gpurun_out/r2_${W}_${N}gpu.json
  python - <<PY
import json
try:
    p=json.load(open('gpurun_out/r2_${W}_${N}gpu.json'))
    print('$W', p['n_gpus'], 'value %.3f G/s' % (p['value']/1e9), 'ms/step %.2f' % p['ms_per_step'], p.get('stage_info'), 'e2e %.3f' % (((p.get('e2e') or {}).get('value') or 0)/1e9), p.get('collectives'))
except Exception as e: print('unreadable', e)
PY
done
