"""Oracle ground removal.

percentile mode follows utils/tower_extraction.py:79-93 (base = np.percentile(z, 25); keep
z > base + 3.0; fall back to base + 1.0 when fewer than 1000 points survive) and the
test/main_ground.py:118-131 variant (ground = z < pct10 + 4).  numpy's percentile is called for real.

grid mode (min-z per XY cell, height above ground) is the north_star extension: NO reference
implementation exists -> self-oracle, PARITY UNPINNED by construction.
"""
import numpy as np


def percentile_keep_mask(z_f32, pct=25, offset=3.0, min_keep=1000, fallback_offset=1.0):
    z = np.asarray(z_f32)
    base = np.percentile(z, pct)
    mask = z > (base + offset)
    used = offset
    if int(mask.sum()) < min_keep:
        mask = z > (base + fallback_offset)
        used = fallback_offset
    return mask, base, used


def grid_min_keep_mask(points_f32, cell=2.0, hag=3.0):
    """keep = z - min_z(cell) > hag, cells = floor((xy - min_xy)/cell) in float32 arithmetic."""
    p = np.asarray(points_f32, dtype=np.float32)
    if p.shape[0] == 0:
        return np.zeros(0, bool), np.zeros(0, np.float32)
    c = np.float32(cell)
    mn = p[:, :2].min(axis=0)
    ij = np.floor((p[:, :2] - mn) / c).astype(np.int64)
    ny = int(ij[:, 1].max()) + 1
    cid = ij[:, 0] * ny + ij[:, 1]
    zmin = np.full(int(cid.max()) + 1, np.inf, dtype=np.float32)
    np.minimum.at(zmin, cid, p[:, 2])
    ground_z = zmin[cid]
    return (p[:, 2] - ground_z) > np.float32(hag), ground_z
