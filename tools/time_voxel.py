"""Ad-hoc stage timing of the voxel path (not the bench): python tools/time_voxel.py [n_points]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import synth, device as dv

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
towers = max(2, n // 2_000_000)
t0 = time.time()
rec = synth.corridor_records(n, towers, "flat", 2)
print(f"generated {n} pts in {time.time()-t0:.1f}s", flush=True)
dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
torch.cuda.synchronize()
for want in (("mean",), ("lattice", "f32")):
    for it in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dv.voxel_downsample(dl, 0.1, 500000, want=want)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"want={want} iter {it}: {ms:.3f} ms  -> {n/ms/1e6:.2f} Gpt/s  M={res.count} plan={res.plan}", flush=True)
