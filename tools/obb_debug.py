"""Compare the device hull (pch_obb_batch workspace) with Qhull, cluster by cluster."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from scipy.spatial import ConvexHull
from pointcloudhookup_b200 import _native, device as dv
from oracle import obb
import test_gpu_obb as T

rng = np.random.default_rng(17)
clusters = T._clusters(rng)
lib = _native.lib()
MAXC, MAXF, HASH, STACK = 32768, 16384, 65536, 3 * 16384
al = lambda b: (b + 255) // 256 * 256
sizes = [MAXC * 4, MAXF * 12, MAXF * 32, HASH * 8, HASH * 4, STACK * 16, MAXC * 4, (MAXC // 32 + 1) * 4, MAXF * 12]
offs = np.concatenate([[0], np.cumsum([al(s) for s in sizes])])
for ci, c in enumerate(clusters):
    rows = torch.from_numpy(c).cuda()
    rg = torch.tensor([[0, len(c)]], dtype=torch.int64).cuda()
    out = torch.zeros(144, dtype=torch.uint8).cuda()
    wsb = lib.pch_obb_workspace_bytes(1)
    ws = torch.zeros(wsb, dtype=torch.uint8).cuda()
    _native.check(lib.pch_obb_batch(rows.data_ptr(), rg.data_ptr(), 1, out.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream))
    r = out.cpu().numpy().view(dv.OBB_DTYPE)[0]
    F, V = int(r["n_faces"]), int(r["n_vertices"])
    w = ws.cpu().numpy()
    faces = w[offs[1]: offs[1] + F * 12].view(np.int32).reshape(F, 3)
    planes = w[offs[2]: offs[2] + F * 32].view(np.float64).reshape(F, 4)
    verts = w[offs[6]: offs[6] + V * 4].view(np.int32)
    h = ConvexHull(c.astype(np.float64), qhull_options="QbB Pp Qt")
    qv = set(h.vertices.tolist())
    dvs = set(verts.tolist())
    # how far outside the device hull is any point?
    viol = (c.astype(np.float64) @ planes[:, :3].T - planes[:, 3]).max()
    t, ext, vol = obb.min_volume_box_all_faces(c)
    R = np.asarray(r["rotation"]); ctr = np.asarray(r["center"])
    loc = (c.astype(np.float64) - ctr) @ R
    print("   device axes orthonormal err", np.abs(R.T @ R - np.eye(3)).max(), "true ptp along device axes", np.round(np.ptp(c.astype(np.float64) @ R, axis=0), 4),
          "max |loc| - ext/2", np.round(np.abs(loc).max(0) - np.asarray(r["extents"]) / 2, 4))
    qn = h.equations[:, :3]
    print("   normal (col 2) best |dot| with a qhull normal", np.abs(qn @ R[:, 2]).max(), " col0:", np.abs(qn @ R[:, 0]).max(), " col1:", np.abs(qn @ R[:, 1]).max(),
          " device planes vs col2:", np.abs(planes[:, :3] @ R[:, 2]).max())
    print(f"cluster {ci}: n={len(c)} status={r['status']} cand={r['n_candidates']} F={F} V={V} qhull F={len(h.simplices)} V={len(qv)} "
          f"missing={len(qv - dvs)} extra={len(dvs - qv)} max plane violation={viol:.3e} vol dev={r['volume']:.6f} oracle={vol:.6f} "
          f"ext dev={np.round(r['extents'],4)} oracle={np.round(ext,4)}", flush=True)
