"""Small end-to-end pass over every kernel, meant to run under compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import synth, device as dv, towers as tw, geo, pipeline

n = 60_000
rec = synth.corridor_records(n, 2, "hilly", 5, (0.86, 0.085, 0.005, 0.05))
dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
v = dv.voxel_downsample(dl, 0.1, 17_001, want=("mean", "lattice", "f32"))
print("voxels", v.count)
recs, mm = dv.encode_records(v.lattice, 34)
x = dv.decode_xyz(dl, torch.float64); y = dv.decode_xyz(dl, torch.float32)
q = dv.quantise(v.mean, synth.SCALES, synth.OFFSETS)
pv = dv.voxel_downsample_points(x[:5000].contiguous(), 0.25)
st = tw.run_stages(v.f32)
print("G", st.filtered.shape[0], "K", st.n_clusters)
st2 = tw.run_stages(v.f32, ground="grid")
res = pipeline.run_pipeline(dl, 0.1, 20_000)
print("towers", len(res.towers))
lat = np.linspace(-90, 90, 721); lon = -180 + 0.25 * np.arange(1440)
g = (30 * np.sin(np.radians(lat))[:, None] * np.cos(np.radians(lon))[None, :]).astype(np.float32)
dg = geo.upload_grid(geo.HostGrid(-90.0, -180.0, 0.25, 0.25, g))
out = geo.las_to_geodetic(dl, dg, -1.0, geo.EPSG4547)
out2 = geo.las_to_geodetic(dl, dg, -1.0, geo.EPSG4547, window=None)
lo, la = geo.gk_inverse([437587.898], [3140691.58])
h = geo.geoid_shift(dg, la, lo, [100.0], 1.0)
torch.cuda.synchronize()
print("ok", float(out[:, 2].sum()), float(h[0]))
