"""Synthetic clouds for the tiled-RANSAC ground tests (shared by the CPU oracle test and the GPU parity test)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from variant_inputs import terrain_cloud  # noqa: E402,F401  (the seeded cloud the golden run used)


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


def tile_sizes(points, tile_size):
    """{tile number: points in the tile}, number of tiles — the reference's edges and masks (test/main_ground.py:84-99)."""
    min_xy = np.min(points[:, :2], axis=0)
    max_xy = np.max(points[:, :2], axis=0)
    x_edges = np.arange(min_xy[0], max_xy[0], tile_size)
    y_edges = np.arange(min_xy[1], max_xy[1], tile_size)
    sizes = {}
    for i in range(len(x_edges) - 1):
        for j in range(len(y_edges) - 1):
            m = (points[:, 0] >= x_edges[i]) & (points[:, 0] < x_edges[i + 1]) & \
                (points[:, 1] >= y_edges[j]) & (points[:, 1] < y_edges[j + 1])
            sizes[i * (len(y_edges) - 1) + j] = int(m.sum())
    return sizes, max(0, len(x_edges) - 1) * max(0, len(y_edges) - 1)


def golden_run_triples(points, case):
    """The draws of the golden run (tests/golden/make_golden_variants.py seeds scikit-learn's estimator with 1000 + call
    number; calls happen for the tiles with >= 10 points in tile order) as a (n_tiles, max_iterations, 3) array."""
    sizes, n_tiles = tile_sizes(points, case["tile_size"])
    fitted = [t for t in sorted(sizes) if sizes[t] >= 10]
    assert [sizes[t] for t in fitted] == case["fit_sizes"]
    seed = {t: 1000 + c for c, t in enumerate(fitted)}
    return replay_triples({t: sizes[t] for t in fitted}, n_tiles, case["max_iterations"], lambda t: seed[t])
