"""Where does the step time go outside kernels?  Times each python-level call of the pipeline with a
device sync after it (so phases don't overlap) and compares the sum with the free-running step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import synth, device as dv, towers as tw, pipeline

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
synth.corridor_records(n, max(2, n // 2_000_000), "hilly", 3, out=pinned.numpy())
dl = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS)
for _ in range(2):
    pipeline.run_pipeline(dl, 0.1, 500000)
torch.cuda.synchronize()

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:28s} host-return {1e3*(t1-t0):7.3f} ms   +drain {1e3*(t2-t1):7.3f} ms", flush=True)
    return r

for rep in range(2):
    t0 = time.perf_counter()
    v = timed("voxel_downsample", lambda: dv.voxel_downsample(dl, 0.1, 500000, want=("f32",)))
    raw = v.f32
    cen = timed("f32_centroid", lambda: dv.f32_centroid(raw))[0]
    zs = timed("f32_shift", lambda: dv.f32_shift(raw, cen, want_z=True))[0]
    r0, r1, gamma = tw.percentile_ranks_f32(raw.shape[0], 25)
    two = timed("select_f32 + D2H", lambda: dv.select_f32(zs, r0, r1).cpu().numpy())
    base = tw.percentile_lerp_f32(two[0], two[1], gamma)
    filt = timed("compact_points", lambda: dv.compact_points(raw, zs, float(base + 3.0), cen))[0]
    db = timed("dbscan_chunked", lambda: dv.dbscan_chunked(filt, 8.0, 80, 50000))
    st = tw.TowerStages(raw, cen.cpu().numpy(), base, 3.0, filt, db.labels, db.n_clusters, db.stats)
    tow = timed("select_towers (host)", lambda: tw.select_towers(st, box="aabb", want_points=False))
    torch.cuda.synchronize()
    print(f"sum of phases (serialised) {1e3*(time.perf_counter()-t0):.3f} ms; towers {len(tow)}")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipeline.run_pipeline(dl, 0.1, 500000)
    torch.cuda.synchronize()
    print(f"free-running run_pipeline {1e3*(time.perf_counter()-t0):.3f} ms")
