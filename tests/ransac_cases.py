"""Synthetic clouds for the tiled-RANSAC ground tests (shared by the CPU oracle test and the GPU parity test)."""
import numpy as np


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


def literal_reference(points, tile_size, distance_threshold, max_iterations, seed_of_tile):
    """test/main_ground.py:77-115 with the real scikit-learn estimator (:8-32), random_state fixed per tile so that the
    draws can be replayed.  Returns (non_ground, ground, {tile number: n points})."""
    from sklearn.linear_model import RANSACRegressor
    min_xy = np.min(points[:, :2], axis=0)
    max_xy = np.max(points[:, :2], axis=0)
    x_edges = np.arange(min_xy[0], max_xy[0], tile_size)
    y_edges = np.arange(min_xy[1], max_xy[1], tile_size)
    non_ground_list, ground_list, sizes = [], [], {}
    for i in range(len(x_edges) - 1):
        for j in range(len(y_edges) - 1):
            tile_mask = (points[:, 0] >= x_edges[i]) & (points[:, 0] < x_edges[i + 1]) & \
                        (points[:, 1] >= y_edges[j]) & (points[:, 1] < y_edges[j + 1])
            tile_points = points[tile_mask]
            t = i * (len(y_edges) - 1) + j
            sizes[t] = len(tile_points)
            if len(tile_points) < 10:
                continue
            ransac = RANSACRegressor(residual_threshold=distance_threshold, max_trials=max_iterations,
                                     random_state=seed_of_tile(t))
            ransac.fit(tile_points[:, :2], tile_points[:, 2])
            inlier_mask = ransac.inlier_mask_
            ground_list.append(tile_points[inlier_mask])
            non_ground_list.append(tile_points[~inlier_mask])
    non_ground = np.vstack(non_ground_list) if non_ground_list else np.zeros((0, 3))
    ground = np.vstack(ground_list) if ground_list else np.zeros((0, 3))
    return non_ground, ground, sizes, (len(x_edges) - 1) * (len(y_edges) - 1)


def replay_triples(sizes, n_tiles, max_iterations, seed_of_tile):
    """(n_tiles, max_iterations, 3) int32: the subsets scikit-learn draws in each tile for random_state=seed_of_tile(t)."""
    from oracle import ransac as orz
    tri = np.zeros((n_tiles, max_iterations, 3), dtype=np.int32)
    for t, n in sizes.items():
        if n >= 10:
            tri[t] = orz.sklearn_triples(n, max_iterations, seed_of_tile(t))
    return tri
