"""GPU: tiled RANSAC ground removal (pch_ransac_*; reference test/main_ground.py:77-115 + :8-32).

Two anchors, both bit-exact on the returned rows:
  * the draws of the product's counter-based generator -> oracle.ransac.remove_ground_tiled_ransac (numpy restatement);
  * the draws scikit-learn itself makes for random_state = 1000 + tile, handed in as `triples=` -> the literal reference
    loops around the REAL sklearn RANSACRegressor (tests/ransac_cases.py::literal_reference).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tiled_ransac_counter_draws_equal_the_oracle(cuda_device):
    import ransac_cases as rc
    from oracle import ransac as orz
    from pointcloudhookup_b200 import ground_ransac as gr
    for seed, origin in ((1, (500000.0, 3.2e6)), (2, (-23.7, -11.3))):        # LAS-like coordinates / a frame around zero
        pts = rc.terrain_cloud(seed, origin=origin)
        exp_ng, exp_g, per_tile = orz.remove_ground_tiled_ransac(pts, 10.0, 0.1, 150, seed=77)
        ng, g, tiles = gr.remove_ground_tiled_ransac(pts, tile_size=10.0, distance_threshold=0.1, max_iterations=150, seed=77,
                                                     return_tiles=True)
        assert g.shape == exp_g.shape and ng.shape == exp_ng.shape, (g.shape, exp_g.shape, ng.shape, exp_ng.shape)
        assert np.array_equal(g, exp_g) and np.array_equal(ng, exp_ng)
        for t, n, info in per_tile:
            assert tiles["status"][t] == 0 and tiles["n_points"][t] == n
            assert tiles["n_trials"][t] == info["n_trials"] and tiles["n_inliers"][t] == info["n_inliers"]
        assert (tiles["status"] == 1).sum() >= 1                               # the thinned corner tile is in neither output
        assert len(g) + len(ng) < len(pts)                                     # ... and so is the strip beyond the last edge
        assert 0.45 < len(g) / (len(g) + len(ng)) < 0.8                        # ~65 % of the points are terrain


def test_tiled_ransac_replaying_sklearns_draws_equals_sklearn(cuda_device):
    import ransac_cases as rc
    from pointcloudhookup_b200 import ground_ransac as gr
    pts = rc.terrain_cloud(5, nx_m=38.0, ny_m=27.0)
    seed_of_tile = lambda t: 1000 + t
    T = 120
    exp_ng, exp_g, sizes, n_tiles = rc.literal_reference(pts, 10.0, 0.1, T, seed_of_tile)
    tri = rc.replay_triples(sizes, n_tiles, T, seed_of_tile)
    ng, g = gr.remove_ground_tiled_ransac(pts, tile_size=10.0, distance_threshold=0.1, max_iterations=T, triples=tri)
    assert np.array_equal(g, exp_g) and np.array_equal(ng, exp_ng)
    assert len(g) > 1000 and len(ng) > 1000


def test_tiled_ransac_reproduces_the_golden_run_of_the_reference_functions(cuda_device):
    """tests/golden/variants_run.json["ransac"]: the reference's own two functions exec'd around the real estimator (see
    make_golden_variants.py); nothing of /root/reference is read here."""
    import json
    import os
    import ransac_cases as rc
    from variant_inputs import digest, terrain_cloud
    from pointcloudhookup_b200 import ground_ransac as gr
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "variants_run.json"), encoding="utf-8"))
    for case in gold["ransac"]:
        pts = terrain_cloud(case["seed"], **{k: tuple(v) if isinstance(v, list) else v for k, v in case["cloud"].items()})
        assert digest(pts) == case["points_sha256"]
        tri = rc.golden_run_triples(pts, case)
        ng, g = gr.remove_ground_tiled_ransac(pts, tile_size=case["tile_size"], distance_threshold=case["distance_threshold"],
                                              max_iterations=case["max_iterations"], triples=tri)
        assert len(g) == case["ground"]["rows"] and digest(g) == case["ground"]["sha256"]
        assert len(ng) == case["non_ground"]["rows"] and digest(ng) == case["non_ground"]["sha256"]


def test_tiled_ransac_edge_cases(cuda_device):
    import torch
    from pointcloudhookup_b200 import ground_ransac as gr
    rng = np.random.default_rng(0)
    small = np.column_stack([rng.uniform(0, 8, (500, 2)), rng.normal(0, 0.01, 500)])
    ng, g = gr.remove_ground_tiled_ransac(small, tile_size=10.0)               # extent < tile: np.arange gives one edge, no tile
    assert ng.shape == (0, 3) and g.shape == (0, 3)
    with pytest.raises(ValueError):
        gr.remove_ground_tiled_ransac(np.zeros((0, 3)))
    # device in, device out; a perfectly planar tile (every point an inlier: the trial budget collapses to 1)
    xy = rng.uniform(0, 25, (4000, 2))
    flat = np.column_stack([xy, 3.0 + 0.5 * xy[:, 0] - 0.25 * xy[:, 1]])
    ng, g, tiles = gr.remove_ground_tiled_ransac(torch.from_numpy(flat).to(cuda_device), tile_size=10.0, distance_threshold=0.1,
                                                 return_tiles=True)
    assert isinstance(g, torch.Tensor) and g.is_cuda and ng.shape[0] == 0
    assert np.all(tiles["n_trials"][tiles["status"] == 0] == 1)
    inside = (xy[:, 0] < xy[:, 0].min() + 20) & (xy[:, 1] < xy[:, 1].min() + 20)
    assert g.shape[0] == int(inside.sum())
