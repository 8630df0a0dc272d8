"""Seeded inputs shared by make_golden_variants.py (which runs the reference's source lines on them) and the tests
(which run the oracle / the product on them)."""
import hashlib

import numpy as np


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def crop_inputs(tower_data, seed):
    """Synthetic float64 cloud around the reference's four logged towers (same recipe in the tests)."""
    rng = np.random.default_rng(seed)
    parts = []
    for t in tower_data:
        c = np.array([t["x"], t["y"], t["z"]])
        parts.append(c + rng.uniform(-1.5, 2.5, (6000, 3)) * np.array([t["width"], t["width"], t["height"]]))
    parts.append(np.array([437500.0, 3140200.0, 90.0]) + rng.uniform(-600, 600, (20000, 3)) * np.array([1, 1, 0.1]))
    pts = np.concatenate(parts)
    return np.round(pts[rng.permutation(len(pts))], 3)


def merge_inputs(seed):
    rng = np.random.default_rng(seed)
    K = int(rng.integers(20, 70))
    centres = rng.uniform(0, 80, (K, 3)) * np.array([1.0, 1.0, 0.25])
    sizes = rng.integers(3, 300, K)
    pts, labs = [], []
    for k in range(K):
        pts.append((centres[k] + rng.normal(0, 0.5, (sizes[k], 3))).astype(np.float32))
        labs.append(np.full(sizes[k], k, dtype=np.int32))
    pts.append(rng.uniform(0, 80, (100, 3)).astype(np.float32))
    labs.append(np.full(100, -1, dtype=np.int32))
    pts, labs = np.concatenate(pts), np.concatenate(labs)
    perm = rng.permutation(len(labs))
    return pts[perm], labs[perm]


def terrain_cloud(seed, nx_m=47.0, ny_m=33.0, per_m2=18.0, origin=(500000.0, 3.2e6), outliers=0.35, noise=0.03):
    """A gently rolling surface sampled at random xy, vegetation / structure points above it, and a sparse corner
    (a tile with fewer than 10 points).  Absolute projected coordinates like a LAS file's."""
    rng = np.random.default_rng(seed)
    n = int(nx_m * ny_m * per_m2)
    xy = rng.uniform(0, 1, size=(n, 2)) * np.array([nx_m, ny_m])
    z = 120.0 + 0.04 * xy[:, 0] - 0.03 * xy[:, 1] + 0.4 * np.sin(xy[:, 0] / 9.0) + rng.normal(0, noise, n)
    k = int(outliers * n)
    pick = rng.choice(n, size=k, replace=False)
    z[pick] += rng.uniform(0.3, 30.0, k)
    sparse = (xy[:, 0] < 10.0) & (xy[:, 1] < 10.0)            # thin the first tile down to a handful of points
    keep = ~sparse | (rng.uniform(size=n) < 4.0 / max(1, int(sparse.sum())))
    pts = np.column_stack([xy + np.asarray(origin), z])[keep]
    return np.ascontiguousarray(pts)
