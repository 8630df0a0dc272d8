"""Oracle-backed stand-in for tiles.DeviceClusterer (test infrastructure; CPU only): the same interface on host
tensors, with the real scikit-learn DBSCAN for the core phase and a literal restatement of scikit-learn's
border rule (a border point joins the first cluster that reaches it = the smallest cluster id among the
clusters that own a core point within eps) for the finish phase.  Lets the world-size-2 gloo test and the
in-process thread test run the whole halo protocol of tiles.tile_dbscan without a GPU."""
import math

import numpy as np
import torch
from sklearn.cluster import DBSCAN
from sklearn.neighbors import NearestNeighbors

from pointcloudhookup_b200.tiles import I64_MAX, STATS_DTYPE


class OracleClusterer:
    def extent(self, P, axis):
        if P.shape[0] == 0:
            return math.inf, -math.inf
        p = P.numpy().astype(np.float64)
        s = p[:, 0] * axis[0] + p[:, 1] * axis[1]
        return float(s.min()), float(s.max())

    def band_pair(self, P, axis, hi_left, lo_right):
        p = P.numpy().astype(np.float64)
        s = p[:, 0] * axis[0] + p[:, 1] * axis[1]
        sel = lambda m: torch.from_numpy(np.nonzero(m)[0].astype(np.int32))
        empty = torch.zeros(0, dtype=torch.int32)
        return (sel(s <= hi_left) if hi_left is not None and len(p) else empty,
                sel(s >= lo_right) if lo_right is not None and len(p) else empty)

    def cores(self, P, eps, min_samples):
        pts = P.numpy()
        db = DBSCAN(eps=eps, min_samples=min_samples, algorithm="ball_tree").fit(pts)
        core = np.zeros(len(pts), bool)
        core[db.core_sample_indices_] = True
        lab = np.where(core, db.labels_, -1).astype(np.int32)
        self._state = (pts, float(eps), core, lab)
        return torch.from_numpy(lab), int(db.labels_.max()) + 1 if len(pts) else 0

    def shared_report(self, labels, pos, extra, lo, hi, base, k):
        t = np.full(k, I64_MAX, dtype=np.int64)
        lab = labels.numpy()
        own = lab[lo:hi]
        ok = own >= 0
        np.minimum.at(t, own[ok], base + np.nonzero(ok)[0].astype(np.int64))
        at = lab[pos.numpy().astype(np.int64)] if pos.numel() else np.zeros(0, np.int32)
        return at.astype(np.int32), extra.numpy().astype(np.int64), t

    def finish(self, label_map, n_global, own_lo, own_hi):
        pts, eps, core, lab = self._state
        out = np.full(len(pts), -1, dtype=np.int32)
        m = np.asarray(label_map, dtype=np.int32)
        out[core] = m[lab[core]]
        nn = NearestNeighbors(radius=eps, algorithm="ball_tree").fit(pts)
        non_core = np.nonzero(~core)[0]
        if len(non_core):
            neigh = nn.radius_neighbors(pts[non_core], return_distance=False)
            for i, nb in zip(non_core, neigh):
                cand = out[nb[core[nb]]]
                cand = cand[cand >= 0]
                if len(cand):
                    out[i] = cand.min()
        stats = np.zeros(n_global, dtype=STATS_DTYPE)
        stats["min"] = np.inf
        stats["max"] = -np.inf
        own = np.arange(own_lo, own_hi)
        l = out[own]
        ok = l >= 0
        np.add.at(stats["count"], l[ok], 1)
        np.add.at(stats["sum"], l[ok], pts[own][ok].astype(np.float64))
        np.minimum.at(stats["min"], l[ok], pts[own][ok])
        np.maximum.at(stats["max"], l[ok], pts[own][ok])
        return torch.from_numpy(out), stats


def corridor_candidates(seed, n_tiles, per_tile=3000, tile_len=120.0):
    """Small candidate clouds along the x axis: blobs (towers), a few of them centred ON the cuts, a thin line of
    points crossing every cut (a conductor), background noise.  Returns the per-tile (n,3) float32 arrays; tile t
    covers x in [t*tile_len, (t+1)*tile_len)."""
    rng = np.random.default_rng(seed)
    total = n_tiles * tile_len
    pts = [rng.uniform([0, -25, 0], [total, 25, 6], (n_tiles * per_tile // 3, 3))]          # sparse background
    for t in range(n_tiles):
        for cx in (t * tile_len + 30.0, t * tile_len + 75.0, (t + 1) * tile_len - 1.5):      # the last one straddles the cut
            k = per_tile // 5
            pts.append(np.column_stack([rng.normal(cx, 3.0, k), rng.normal(0, 3.0, k), rng.uniform(0, 35, k)]))
    line = np.column_stack([np.linspace(0, total, n_tiles * 1500), np.full(n_tiles * 1500, 12.0), np.full(n_tiles * 1500, 30.0)])
    pts.append(line + rng.normal(0, 0.2, line.shape))
    allp = np.concatenate(pts).astype(np.float32)
    allp = allp[(allp[:, 0] >= 0) & (allp[:, 0] < total)]
    tile = np.minimum((allp[:, 0].astype(np.float64) // tile_len).astype(int), n_tiles - 1)
    out = []
    for t in range(n_tiles):
        p = allp[tile == t]
        out.append(p[rng.permutation(len(p))])                                               # no spatial order inside a tile
    return out
