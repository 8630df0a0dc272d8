"""TEST INFRASTRUCTURE (CPU oracle) — per-tower box crop and preview subsample restated in numpy.

crop: test/kuangxuan.py:58-79, literally (one boolean mask over all points per tower).
sample: the product's keyed bijection (pch_sample_indices) restated; the reference itself calls
np.random.choice(len(xyz), k, replace=False) on the unseeded global RNG (pyGUI_towers_test.py:174-177;
ui/vtk_widget.py:115-118), which no implementation can reproduce — parity for the draw is unpinned by
construction; what is checked is "k distinct in-range indices" plus this restatement.
"""
import numpy as np


def kuangxuan_bounds(tower):
    """test/kuangxuan.py:63-71."""
    w = tower['width']
    original_h = tower['height']
    cx, cy, cz = tower['x'], tower['y'], tower['z']
    x_min, x_max = cx - w / 1, cx + w / 0.6
    y_min, y_max = cy - w / 2, cy + w / 1
    z_min, z_max = cz - original_h / 1, cz + original_h * 2
    return np.array([x_min, y_min, z_min, x_max, y_max, z_max])


def crop_boxes(points, boxes):
    """test/kuangxuan.py:73-79 per box."""
    out = []
    for x_min, y_min, z_min, x_max, y_max, z_max in np.asarray(boxes, dtype=np.float64).reshape(-1, 6):
        mask = ((points[:, 0] >= x_min) & (points[:, 0] <= x_max) &
                (points[:, 1] >= y_min) & (points[:, 1] <= y_max) &
                (points[:, 2] >= z_min) & (points[:, 2] <= z_max))
        out.append(points[mask])
    return out


def _mix32(h):
    h = h.astype(np.uint64)
    M = np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85ebca6b)) & M
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xc2b2ae35)) & M
    h ^= h >> np.uint64(16)
    return h


def sample_indices(n, k, seed):
    j = np.arange(k, dtype=np.uint64)
    if not seed:
        return (j * np.uint64(n)) // np.uint64(k)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    bits = 2
    while (1 << bits) < n:
        bits += 2
    hb = np.uint64(bits // 2)
    mask = np.uint64((1 << (bits // 2)) - 1)
    v = j.copy()
    todo = np.ones(k, dtype=bool)
    while todo.any():
        l = (v[todo] >> hb) & mask
        r = v[todo] & mask
        for rnd in range(4):
            key = np.uint64(((seed >> (16 * (rnd & 1))) ^ (0x9e3779b9 * (rnd + 1)) ^ (seed >> 32)) & 0xFFFFFFFF)
            f = _mix32(r ^ key) & mask
            l, r = r, l ^ f
        v[todo] = (l << hb) | r
        todo = v >= np.uint64(n)
    return v
