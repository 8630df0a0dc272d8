"""Timing of the tiled RANSAC ground removal on a synthetic corridor, and of scikit-learn's RANSACRegressor on a sample
of its tiles (the reference's per-tile cost): python tools/prof_ransac.py [n_points] [json_out]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudhookup_b200 import _native, ground_ransac as gr

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
out_path = sys.argv[2] if len(sys.argv) > 2 else None
L, W = 2000.0, 500.0                                   # corridor, metres: 20 points / m^2 at 20 M points
g = torch.Generator(device="cuda").manual_seed(3)
xy = torch.rand((n, 2), generator=g, device="cuda", dtype=torch.float64) * torch.tensor([L, W], device="cuda", dtype=torch.float64)
z = 120.0 + 0.04 * xy[:, 0] - 0.03 * xy[:, 1] + 0.4 * torch.sin(xy[:, 0] / 9.0) + 0.03 * torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
up = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
z = z + torch.where(up < 0.35, 0.3 + 85.0 * up, torch.zeros_like(up))     # 35 % of the points above the surface
P = torch.cat([xy + torch.tensor([500000.0, 3.2e6], device="cuda", dtype=torch.float64), z[:, None]], dim=1).contiguous()
del xy, z, up
lib = _native.lib()
rows = []
for rep in range(3):
    lib.pch_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ng, gnd, tiles = gr.remove_ground_tiled_ransac(P, tile_size=10.0, distance_threshold=0.1, max_iterations=1000, seed=rep, return_tiles=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    lib.pch_profile_enable(0)
    buf = ctypes.create_string_buffer(65536); lib.pch_profile_report(buf, 65536)
    ok = tiles["status"] == 0
    rows.append({"wall_ms": 1e3 * (t1 - t0), "kernels": buf.value.decode().strip(), "ground": int(gnd.shape[0]), "non_ground": int(ng.shape[0]),
                 "tiles": int(ok.sum()), "trials_mean": float(tiles["n_trials"][ok].mean()), "trials_max": int(tiles["n_trials"][ok].max()),
                 "points_per_tile_median": int(np.median(tiles["n_points"][ok]))})
    print(rows[-1], flush=True)
# the reference's estimator on a few of the same tiles (host, one core like the reference's loop)
from sklearn.linear_model import RANSACRegressor
host = P[: min(n, 4_000_000)].cpu().numpy()
x0, y0 = host[:, 0].min(), host[:, 1].min()
t_fit, n_fit, k = 0.0, 0, 0
for i in range(40):
    m = (host[:, 0] >= x0 + 10 * i) & (host[:, 0] < x0 + 10 * (i + 1)) & (host[:, 1] >= y0 + 10 * (i % 7)) & (host[:, 1] < y0 + 10 * (i % 7 + 1))
    tp = host[m]
    if len(tp) < 10:
        continue
    t0 = time.perf_counter()
    RANSACRegressor(residual_threshold=0.1, max_trials=1000).fit(tp[:, :2], tp[:, 2])
    t_fit += time.perf_counter() - t0
    n_fit += len(tp); k += 1
cpu = {"tiles": k, "points": n_fit, "fit_seconds": t_fit, "points_per_s": n_fit / max(t_fit, 1e-9)}
print(cpu, flush=True)
res = {"n": n, "runs": rows, "sklearn_fit_only": cpu, "gpu_points_per_s": n / (min(r["wall_ms"] for r in rows) * 1e-3)}
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
