"""Oracle for the tiled RANSAC ground removal (TEST INFRASTRUCTURE — never imported by the product).

Follows test/main_ground.py:77-115 (`remove_ground_tiled_ransac`: np.arange tile edges, the two tile loops,
`len(tile_points) < 10: continue`, np.vstack of the per-tile results) and :8-32 (`remove_ground_ransac`:
RANSACRegressor(residual_threshold, max_trials) on X = xy, y = z; ground = inlier_mask_).

The estimator itself is scikit-learn (third party, not under /root/reference; 1.9.0 in this image):
`ransac_inlier_mask` restates the trial loop of sklearn/linear_model/_ransac.py::RANSACRegressor.fit for its
defaults (LinearRegression, min_samples = 3, loss = absolute error, max_skips / stop_n_inliers / stop_score =
inf, stop_probability = 0.99) with the 3-point LinearRegression solved in closed form.  It is pinned against
the REAL RANSACRegressor in tests/test_cpu_oracle_golden.py by replaying the triples scikit-learn's own
`sample_without_replacement` draws from RandomState(seed) (`sklearn_triples`): identical inlier masks.
The reference passes no random_state, so its own output is not reproducible run to run; parity is defined on
the algorithm given the same draws.
"""
import numpy as np

MASK64 = (1 << 64) - 1
MIN_TILE_POINTS = 10          # test/main_ground.py:101


def mix64(z):
    """splitmix64 finaliser on Python ints (the product's counter-based generator)."""
    z = (z + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def counter_triple(seed, tile, trial, n):
    """Three distinct row numbers in [0, n) for trial `trial` (1-based) of tile `tile`."""
    s = mix64((mix64(mix64(seed & MASK64) ^ tile) + trial) & MASK64)
    s = mix64(s)
    i1 = s % n
    while True:
        s = mix64(s)
        i2 = s % n
        if i2 != i1:
            break
    while True:
        s = mix64(s)
        i3 = s % n
        if i3 != i1 and i3 != i2:
            break
    return int(i1), int(i2), int(i3)


def sklearn_triples(n, max_trials, seed):
    """The subsets RANSACRegressor(random_state=seed) draws for n samples: its only consumer of the generator is
    sample_without_replacement(n_samples, min_samples, random_state) once per trial."""
    from sklearn.utils import check_random_state
    from sklearn.utils.random import sample_without_replacement
    rs = check_random_state(seed)
    return np.array([sample_without_replacement(n, 3, random_state=rs) for _ in range(max_trials)], dtype=np.int32)


def dynamic_max_trials(n_inliers, n_samples, min_samples=3, probability=0.99):
    """sklearn/linear_model/_ransac.py::_dynamic_max_trials."""
    eps = np.spacing(1)
    inlier_ratio = n_inliers / float(n_samples)
    nom = max(eps, 1 - probability)
    denom = max(eps, 1 - inlier_ratio ** min_samples)
    if nom == 1:
        return 0
    if denom == 1:
        return float("inf")
    return abs(float(np.ceil(np.log(nom) / np.log(denom))))


def plane_through(p1, p2, p3):
    """(a, b) of z - z1 = a (x - x1) + b (y - y1) through three points, or None when they are collinear in xy."""
    ux, uy, uz = p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]
    vx, vy, vz = p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]
    det = ux * vy - uy * vx
    if det == 0.0 or not np.isfinite(det):
        return None
    return (uz * vy - uy * vz) / det, (ux * vz - uz * vx) / det


def residuals(points, anchor, a, b):
    t1 = a * (points[:, 0] - anchor[0])
    t2 = b * (points[:, 1] - anchor[1])
    return np.abs((points[:, 2] - anchor[2]) - (t1 + t2))


def ransac_inlier_mask(points, residual_threshold=0.1, max_trials=1000, triple_of_trial=None, stop_probability=0.99):
    """RANSACRegressor(...).fit(points[:, :2], points[:, 2]).inlier_mask_ for the draws `triple_of_trial(k)`
    (k = 1, 2, ...).  Returns (mask, info) or (None, info) when no consensus set was found."""
    pts = np.asarray(points, dtype=np.float64)
    n = pts.shape[0]
    n_best, score_best, mask_best, plane_best = 1, -np.inf, None, None
    trials, limit = 0, float(max_trials)
    while trials < limit:
        trials += 1
        i1, i2, i3 = triple_of_trial(trials)
        p1 = pts[i1]
        ab = plane_through(np.float64(p1), np.float64(pts[i2]), np.float64(pts[i3]))
        if ab is None:
            continue
        res = residuals(pts, p1, ab[0], ab[1])
        mask = res <= residual_threshold
        cnt = int(mask.sum())
        if cnt < n_best:
            continue
        zz = pts[mask, 2] - p1[2]
        ss_res = float(np.sum(res[mask] ** 2))
        ss_tot = float(np.sum((zz - zz.mean()) ** 2))
        score = 1.0 - ss_res / ss_tot if ss_tot > 0.0 else (1.0 if ss_res == 0.0 else 0.0)
        if cnt == n_best and score < score_best:
            continue
        n_best, score_best, mask_best, plane_best = cnt, score, mask, (p1.copy(), ab[0], ab[1])
        limit = min(limit, dynamic_max_trials(n_best, n, 3, stop_probability))
    return mask_best, {"n_trials": trials, "n_inliers": n_best if mask_best is not None else 0, "score": score_best,
                       "plane": plane_best}


def tile_edges(points, tile_size):
    """x_edges, y_edges of test/main_ground.py:84-91."""
    min_xy = np.min(points[:, :2], axis=0)
    max_xy = np.max(points[:, :2], axis=0)
    return np.arange(min_xy[0], max_xy[0], tile_size), np.arange(min_xy[1], max_xy[1], tile_size)


def remove_ground_tiled_ransac(points, tile_size=10.0, distance_threshold=0.1, max_iterations=1000, seed=0, triples=None):
    """The reference's loops, literally; the estimator replaced by `ransac_inlier_mask`.
    triples: optional dict tile_number -> (max_iterations, 3) int array of draws (tile_number = i*(len(y_edges)-1)+j);
    otherwise the product's counter-based generator keyed by (seed, tile_number, trial).
    Returns (non_ground, ground, per_tile) with per_tile = list of (tile_number, n_points, info)."""
    points = np.asarray(points, dtype=np.float64)
    x_edges, y_edges = tile_edges(points, tile_size)
    non_ground_list, ground_list, per_tile = [], [], []
    nty = len(y_edges) - 1
    for i in range(len(x_edges) - 1):
        for j in range(len(y_edges) - 1):
            tile_mask = (points[:, 0] >= x_edges[i]) & (points[:, 0] < x_edges[i + 1]) & \
                        (points[:, 1] >= y_edges[j]) & (points[:, 1] < y_edges[j + 1])
            tile_points = points[tile_mask]
            if len(tile_points) < MIN_TILE_POINTS:
                continue
            t = i * nty + j
            n = len(tile_points)
            if triples is not None:
                tri = triples[t]
                draw = lambda k, tri=tri: tuple(int(v) for v in tri[k - 1])
            else:
                draw = lambda k, t=t, n=n: counter_triple(seed, t, k, n)
            mask, info = ransac_inlier_mask(tile_points, distance_threshold, max_iterations, draw)
            if mask is None:
                raise ValueError("RANSAC could not find a valid consensus set")
            non_ground_list.append(tile_points[~mask])
            ground_list.append(tile_points[mask])
            per_tile.append((t, n, info))
    non_ground = np.vstack(non_ground_list) if non_ground_list else np.zeros((0, 3))
    ground = np.vstack(ground_list) if ground_list else np.zeros((0, 3))
    return non_ground, ground, per_tile
