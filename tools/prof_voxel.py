"""Short voxel-stage driver for ncu captures: python tools/prof_voxel.py [n_points] [calls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudhookup_b200 import synth, device as dv
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
synth.corridor_records(n, max(2, n // 2_000_000), "hilly", 3, out=pinned.numpy())
dl = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS)
for _ in range(calls):
    r = dv.voxel_downsample(dl, 0.1, 500000, want=("f32", "z32"))
torch.cuda.synchronize()
print("M", r.count)
