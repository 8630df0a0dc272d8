"""Driver for timing / profiling the device OBB and the grid-ground kernels on the bench workload:
python tools/prof_obb.py [n_points]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import synth, device as dv, towers as tw
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
pinned = torch.empty(n * 34, dtype=torch.uint8, pin_memory=True)
synth.corridor_records(n, max(2, n // 2_000_000), "hilly", 3, out=pinned.numpy())
dl = dv.upload_records(pinned, n, 34, synth.SCALES, synth.OFFSETS)
raw = dv.voxel_downsample(dl, 0.1, 500000, want=("f32",)).f32
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    filt, cen, _, _, _ = tw.ground_filter_grid(raw, 2.0, 3.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"grid ground: {1e3*(t1-t0):.2f} ms, kept {filt.shape[0]} of {raw.shape[0]}", flush=True)
stages = tw.run_stages(raw)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    towers = tw.select_towers(stages, box="obb", want_points=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"select_towers(obb): {1e3*(t1-t0):.1f} ms, K={stages.n_clusters}, towers={len(towers)}", flush=True)
# the batch alone, with its statistics
st = stages.stats
K = stages.n_clusters
diag = np.linalg.norm((st["max"][:K] - st["min"][:K]).astype(np.float64), axis=1)
cand = [l for l in range(K) if diag[l] > 15.0 and st["count"][l] >= 4]
rows, off = dv.cluster_major_points(stages.filtered, stages.labels, st["count"][:K])
torch.cuda.synchronize(); t0 = time.perf_counter()
res = dv.obb_batch(rows, np.array([[off[l], off[l + 1]] for l in cand], dtype=np.int64))
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"obb_batch: {len(cand)} clusters in {1e3*(t1-t0):.1f} ms; status counts {np.bincount(res['status'], minlength=4).tolist()}; "
      f"faces median {int(np.median(res['n_faces']))} max {int(res['n_faces'].max())}; candidates median {int(np.median(res['n_candidates']))} "
      f"max {int(res['n_candidates'].max())}; points median {int(np.median(st['count'][cand]))}", flush=True)
nc = res["n_candidates"]
print("candidate quantiles 50/90/99/max:", np.percentile(nc, [50, 90, 99, 100]).astype(int).tolist(),
      " >2048:", int((nc > 2048).sum()), " >8192:", int((nc > 8192).sum()), " faces 50/90/99/max:", np.percentile(res["n_faces"], [50, 90, 99, 100]).astype(int).tolist())
from pointcloudhookup_b200 import obb as hobb
bad = [l for l, r in zip(cand, res) if r["status"] != 0]
t0 = time.perf_counter()
for l in bad:
    pts = rows[int(off[l]): int(off[l + 1])].cpu().numpy()
    try:
        hobb.bounding_box_oriented(pts)
    except Exception as e:
        pass
print(f"host fallback for {len(bad)} clusters: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
# kernel-only time of the batch (library events) against the wall time of the call
import ctypes
from pointcloudhookup_b200 import _native
lib = _native.lib()
rgs = np.array([[off[l], off[l + 1]] for l in cand], dtype=np.int64)
for rep in range(3):
    lib.pch_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = dv.obb_batch(rows, rgs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    lib.pch_profile_enable(0)
    buf = ctypes.create_string_buffer(65536); lib.pch_profile_report(buf, 65536)
    print(f"obb_batch rep {rep}: wall {1e3*(t1-t0):.1f} ms; kernels: {buf.value.decode().strip()}", flush=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
ws = torch.empty(lib.pch_obb_workspace_bytes(len(cand)), dtype=torch.uint8, device=rows.device)
torch.cuda.synchronize(); print(f"workspace alloc {ws.numel()/1e9:.2f} GB: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
