"""CPU: the oracle reproduces what the UNMODIFIED reference modules produced
(tests/golden/reference_run.json, made by tests/golden/make_golden.py)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))
CASES = ["dense_wires", "towers", "hilly"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("golden")


def make_input(case, workdir):
    from pointcloudhookup_b200 import synth
    c = GOLD[case]["config"]
    path = os.path.join(workdir, f"{case}.las")
    if not os.path.exists(path):
        synth.write_corridor_las(path, c["n"], c["towers"], c["terrain"], c["seed"], tuple(c["fractions"]))
    return path, c


@pytest.mark.parametrize("case", CASES)
def test_generator_is_deterministic(case, workdir):
    path, _ = make_input(case, workdir)
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == GOLD[case]["input_sha256"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_voxel_matches_reference_run(case, workdir):
    from oracle import las_io, voxel
    path, c = make_input(case, workdir)
    out = os.path.join(workdir, f"{case}_ds.las")
    voxel.run_voxel_downsampling(path, out, c["voxel"], c["chunk"])
    ds = las_io.read_las(out)
    g = GOLD[case]["downsample"]
    assert ds["n"] == g["count"]
    assert digest(np.stack([ds["X"], ds["Y"], ds["Z"]], 1)) == g["xyz_sha256"]
    pts = np.stack(las_io.scaled(las_io.read_las(path), 0, 5000), 1)
    pc = voxel.voxel_down_sample(pts, c["voxel"])
    assert pc.shape[0] == GOLD[case]["process_chunk"]["count"] and digest(pc) == GOLD[case]["process_chunk"]["sha256"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_towers_match_reference_run(case, workdir):
    from oracle import towers, voxel
    path, c = make_input(case, workdir)
    out = os.path.join(workdir, f"{case}_ds.las")
    if not os.path.exists(out):
        voxel.run_voxel_downsampling(path, out, c["voxel"], c["chunk"])
    inter = {}
    from oracle import las_io
    res = towers.extract_towers_arrays(las_io.read_las(out), box="obb", intermediates=inter)
    g = GOLD[case]["towers"]
    assert len(inter["labels"]) == g["n_filtered"] and digest(inter["labels"]) == g["labels_sha256"]
    assert len(res) == g["count"]
    for a, b in zip(res, g["list"]):
        assert np.allclose(a["center"], b["center"], atol=1e-9)
        assert np.allclose(a["extent"], b["extent"], atol=1e-9)
        assert np.allclose(a["rotation"], b["rotation"], atol=1e-9)
        assert abs(a["north_angle"] - b["north_angle"]) < 1e-9
        assert digest(a["points"]) == b["points_sha256"]


def test_oracle_elevation_matches_reference_run():
    from oracle import geoid
    e = GOLD["elevation"]
    t = np.array(e["towers"])
    assert np.array_equal(geoid.ellipsoid_to_orthometric(None, t[:, 0], t[:, 1], t[:, 2]), e["fallback"])
    crop = geoid.read_gtx(os.path.join(HERE, "golden", "egm96_crop_20N35N_105E120E.gtx"))
    got = geoid.ellipsoid_to_orthometric(crop, t[:, 0], t[:, 1], t[:, 2])
    assert np.allclose(got, e["grid_egm96_plus1"], atol=1e-9)


def test_oracle_crs_known_answers():
    """EPSG:4547 -> 4326 pinned by the reference's own data: tower centres of its recorded run
    (test/kuangxuan.py:29-33) vs the same towers as lat/lon (elevation_conversion.py:148-153)."""
    from oracle import crs
    en = [(437587.898, 3140691.58), (437787.178, 3140006.96), (437908.948, 3139606.82), (437676.583, 3140379.50)]
    ll = [(28.379751, 113.363246), (28.373584, 113.365316), (28.369979, 113.366579), (28.376940, 113.364167)]
    lon, lat = crs.gk_inverse([p[0] for p in en], [p[1] for p in en])
    assert np.allclose(lat, [p[0] for p in ll], atol=6e-7) and np.allclose(lon, [p[1] for p in ll], atol=6e-7)


def test_oracle_geoid_analytic_grid():
    """On a grid sampled from N = 30 sin(lat) cos(lon) (the shipped npz), bilinear error is bounded by
    h^2/8 * |f''| per axis."""
    from oracle import geoid
    lat = np.linspace(-90, 90, 721)
    lon = np.linspace(-180, 179.75, 1440)
    g = {"ll_lat": -90.0, "ll_lon": -180.0, "dlat": 0.25, "dlon": 0.25, "rows": 721, "cols": 1440,
         "grid": (30 * np.sin(np.radians(lat))[:, None] * np.cos(np.radians(lon))[None, :]).astype(np.float32)}
    rng = np.random.default_rng(0)
    la, lo = rng.uniform(-89, 89, 2000), rng.uniform(-180, 180, 2000)
    n = geoid.geoid_height(g, la, lo)
    exact = 30 * np.sin(np.radians(la)) * np.cos(np.radians(lo))
    bound = 2 * 30 * np.radians(0.25) ** 2 / 8 + 1e-5
    assert np.abs(n - exact).max() < bound
    # wrap across the seam and node hits
    assert abs(float(geoid.geoid_height(g, 10.0, 179.9)) - 30 * np.sin(np.radians(10)) * np.cos(np.radians(179.9))) < bound
    assert float(geoid.geoid_height(g, 0.25, 0.5)) == float(g["grid"][361, 722])
    assert np.isnan(geoid.geoid_height(g, 91.0, 0.0))


def test_oracle_obb_properties():
    from oracle import obb
    rng = np.random.default_rng(4)
    box = rng.uniform(-1, 1, (4000, 3)) * np.array([10.0, 3.0, 20.0])
    th = np.radians(33)
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    pts = box @ R.T + np.array([100.0, -50.0, 7.0])
    tr, ext = obb.bounding_box_oriented(pts)
    local = (pts - tr[:3, 3]) @ tr[:3, :3]
    assert np.all(np.abs(local) <= ext / 2 + 1e-6)
    assert np.prod(ext) <= np.prod(np.ptp(pts, axis=0)) + 1e-9
    assert np.allclose(sorted(ext), [6, 20, 40], rtol=0.06)      # one normal per 0.1 rad bin: a few percent off the true box
    _, ext_all, _ = obb.min_volume_box_all_faces(pts)
    assert np.allclose(sorted(ext_all), [6, 20, 40], rtol=0.02) and np.prod(ext_all) <= np.prod(ext) * (1 + 1e-12)


def test_oracle_match_towers_matches_reference_run():
    """utils/table_match_gim.py::match_towers, run unmodified for the golden file."""
    from oracle import match
    m = GOLD["match"]
    pc = [{"center": np.array(t["center"])} for t in m["pc"]]
    matched, conv = match.match_towers(m["gim"], pc, grid=None)
    assert [list(x) for x in matched] == m["matched"]
    for c, ref in zip(conv, m["converted"]):
        assert np.allclose(c, ref["converted_center"], atol=1e-9)
    for i, g in enumerate(m["gim"]):
        for j, c in enumerate(conv):
            assert abs(match.haversine(g["lat"], g["lng"], c[1], c[0]) - m["haversine"][i][j]) < 1e-6


def test_crop_oracle_bounds_and_sample_restatement():
    """oracle/crop.py: the kuangxuan box of the reference's first logged tower (test/kuangxuan.py:29-30, 63-71)
    and the sample restatement being a bijection of [0, n)."""
    from oracle import crop as oc
    t = {"id": 8, "height": 17.4, "width": 20.1, "x": 4.37587898e+05, "y": 3.14069158e+06, "z": 1.31457350e+02}
    b = oc.kuangxuan_bounds(t)
    assert np.allclose(b, [437587.898 - 20.1, 3140691.58 - 10.05, 131.45735 - 17.4,
                           437587.898 + 33.5, 3140691.58 + 20.1, 131.45735 + 34.8], rtol=0, atol=1e-9)
    pts = np.array([[b[0], b[1], b[2]], [b[3], b[4], b[5]], [b[0] - 1e-6, b[1], b[2]], [b[3], b[4], b[5] + 1e-6]])
    assert np.array_equal(oc.crop_boxes(pts, b[None])[0], pts[:2])          # bounds are inclusive
    for n in (1, 2, 5, 1000, 70001):
        v = oc.sample_indices(n, n, 99)
        assert np.array_equal(np.sort(v), np.arange(n, dtype=np.uint64))
        k = max(1, n // 3)
        assert np.unique(oc.sample_indices(n, k, 0)).size == k


def _variants():
    import json
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import variant_inputs as vi
    return json.load(open(os.path.join(HERE, "golden", "variants_run.json"), encoding="utf-8")), vi


def test_crop_oracle_reproduces_the_reference_lines():
    """tests/golden/variants_run.json holds what test/kuangxuan.py:60-79 (the reference's own lines, exec'd by
    make_golden_variants.py) produced for the reference's four logged towers on seeded clouds: oracle/crop.py must
    give the same boxes (exactly) and the same `points[mask]` arrays (sha256)."""
    from oracle import crop as oc
    gold, vi = _variants()
    towers = gold["tower_data"]
    assert [t["id"] for t in towers] == [8, 188, 199, 235]
    for case in gold["crop"]:
        pts = vi.crop_inputs(towers, case["seed"])
        assert vi.digest(pts) == case["points_sha256"]
        boxes = np.stack([oc.kuangxuan_bounds(t) for t in towers])
        got = oc.crop_boxes(pts, boxes)
        for t, g, want in zip(towers, got, case["towers"]):
            assert list(oc.kuangxuan_bounds(t)) == want["bounds"]
            assert len(g) == want["count"] and vi.digest(g) == want["sha256"]


def test_cluster_merge_reproduces_the_reference_lines():
    """The merge block test/tttt.py:93-175 exec'd on seeded clusters: the oracle's restatement AND the product's
    stats-based merge (towers.merge_adjacent_clusters) must give the same merged labels, in the same set() order."""
    from oracle import towers as ot
    from pointcloudhookup_b200 import device as dv, towers as tw
    gold, vi = _variants()
    for case in gold["merge"]:
        pts, labs = vi.merge_inputs(case["seed"])
        assert vi.digest(labs) == case["labels_sha256"] and vi.digest(pts) == case["points_sha256"]
        thr = case["merge_threshold"]
        merged = ot.merge_adjacent_clusters(pts, labs, thr)
        assert vi.digest(np.asarray(merged).astype(np.int64)) == case["merged_sha256"]
        assert [int(v) for v in (set(merged) - {-1})] == case["iteration_order"]
        K = int(labs.max()) + 1
        stats = np.zeros(K, dtype=dv.STATS_DTYPE)
        for k in range(K):
            cp = pts[labs == k]
            stats[k] = (len(cp), cp.min(0), cp.max(0), cp.astype(np.float64).sum(0))
        comp, mstats = tw.merge_adjacent_clusters(stats, K, thr)
        got = np.where(labs >= 0, K + comp[np.maximum(labs, 0)], -1).astype(np.int64)
        assert vi.digest(got) == case["merged_sha256"]
        assert len(mstats) == case["n_merged_clusters"]


def test_geoid_nodata_corners_are_reweighted_like_proj():
    """PROJ vgridshift: nodata corners (-88.8888) are dropped and the remaining weights renormalised; only a cell
    with four nodata corners has no value."""
    from oracle import geoid
    nd = np.float32(geoid.NODATA)
    g = np.array([[1.0, 2.0, nd], [3.0, nd, nd], [5.0, 6.0, 7.0]], dtype=np.float32)
    grid = {"ll_lat": 10.0, "ll_lon": 100.0, "dlat": 1.0, "dlon": 1.0, "rows": 3, "cols": 3, "grid": g}
    fx, fy = 0.25, 0.5
    # cell (0,0): corners 1, 2 (east), 3 (north), nodata (north-east)
    w00, w01, w10 = (1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy
    exp = (w00 * 1.0 + w01 * 2.0 + w10 * 3.0) / (w00 + w01 + w10)
    got = geoid.geoid_height(grid, [10.0 + fy], [100.0 + fx])
    assert got[0] == exp
    # cell (0,1): only the south-west corner (2.0) is valid -> exactly that value
    assert geoid.geoid_height(grid, [10.5], [101.5])[0] == 2.0
    # all four valid elsewhere: plain bilinear; all nodata: NaN
    g2 = g.copy()
    g2[0, 1] = nd
    g2[:, 2] = nd
    g2[1, 1] = nd
    grid2 = dict(grid, grid=g2)
    assert np.isnan(geoid.geoid_height(grid2, [10.5], [101.5])[0])
    full = dict(grid, grid=np.array([[1, 2, 3], [3, 4, 5], [5, 6, 7]], dtype=np.float32))
    assert geoid.geoid_height(full, [10.5], [100.25])[0] == (0.375 * 1 + 0.125 * 2 + 0.375 * 3 + 0.125 * 4)


def test_oracle_all_faces_box_is_never_larger_than_the_thinned_search():
    """oracle.obb.min_volume_box_all_faces evaluates every hull-face normal; trimesh's thinned search (one normal per
    0.1 rad bin, first in Qhull's facet order) can only pick among the same candidates."""
    from oracle import obb
    rng = np.random.default_rng(5)
    for trial in range(6):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        n = int(rng.integers(200, 4000))
        p = (rng.uniform(-1, 1, (n, 3)) * rng.uniform(2, 25, 3)) @ q.T + rng.uniform(-50, 50, 3)
        if trial % 2:
            p = np.concatenate([p, rng.normal(0, 3, (n // 3, 3)) + p.mean(0)])
        p = p.astype(np.float32)
        t_all, ext_all, vol_all = obb.min_volume_box_all_faces(p)
        t_tm, ext_tm = obb.bounding_box_oriented(p)
        assert vol_all <= np.prod(ext_tm) * (1 + 1e-12)
        assert abs(np.prod(ext_all) - vol_all) <= 1e-9 * vol_all
        # the box contains every point
        loc = (p.astype(np.float64) - t_all[:3, 3]) @ t_all[:3, :3]
        assert np.all(np.abs(loc) <= ext_all / 2 + 1e-6)
        assert abs(np.linalg.det(t_all[:3, :3]) - 1.0) < 1e-9


def test_ransac_oracle_replaying_sklearns_draws_equals_sklearn():
    """oracle.ransac restates RANSACRegressor's trial loop with a closed-form 3-point plane; with the subsets
    scikit-learn's own generator draws for random_state=seed it must return scikit-learn's inlier mask and stop after
    the same number of trials — on single tiles, and through the reference's tile loops (test/main_ground.py:77-115)."""
    from sklearn.linear_model import RANSACRegressor
    from oracle import ransac as orz
    import ransac_cases as rc
    for seed in range(12):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(12, 1200))
        xy = rng.uniform(0, 10, (n, 2)) + np.array([500000.0, 3.2e6])
        z = 100 + 0.05 * (xy[:, 0] - 500000) - 0.02 * (xy[:, 1] - 3.2e6) + rng.normal(0, 0.03, n)
        z[: int([0.1, 0.4, 0.7][seed % 3] * n)] += rng.uniform(0.5, 25, int([0.1, 0.4, 0.7][seed % 3] * n))
        pts = np.column_stack([xy, z])
        T = 150
        tri = orz.sklearn_triples(n, T, seed)
        mask, info = orz.ransac_inlier_mask(pts, 0.1, T, lambda k: tuple(int(v) for v in tri[k - 1]))
        sk = RANSACRegressor(residual_threshold=0.1, max_trials=T, random_state=seed).fit(pts[:, :2], pts[:, 2])
        assert np.array_equal(mask, sk.inlier_mask_) and info["n_trials"] == sk.n_trials_
    pts = rc.terrain_cloud(5, nx_m=38.0, ny_m=27.0)
    seed_of_tile = lambda t: 1000 + t
    exp_ng, exp_g, sizes, n_tiles = rc.literal_reference(pts, 10.0, 0.1, 120, seed_of_tile)
    tri = rc.replay_triples(sizes, n_tiles, 120, seed_of_tile)
    ng, g, _ = orz.remove_ground_tiled_ransac(pts, 10.0, 0.1, 120, triples={t: tri[t] for t in range(n_tiles)})
    assert np.array_equal(ng, exp_ng) and np.array_equal(g, exp_g)
    assert min(sizes.values()) < 10 and len(g) + len(ng) < len(pts)       # a skipped tile and the strip beyond the last edge


def test_ransac_counter_generator_draws_three_distinct_rows():
    from oracle import ransac as orz
    for n in (3, 4, 10, 1000):
        for k in range(1, 40):
            t = orz.counter_triple(9, 5, k, n)
            assert len(set(t)) == 3 and all(0 <= v < n for v in t)
    assert orz.counter_triple(9, 5, 1, 1000) != orz.counter_triple(9, 6, 1, 1000) != orz.counter_triple(10, 5, 1, 1000)


def test_tiled_ransac_oracle_reproduces_the_reference_functions():
    """variants_run.json["ransac"]: test/main_ground.py:8-32 and :77-115 exec'd by make_golden_variants.py around the REAL
    scikit-learn RANSACRegressor (seeded per call).  The oracle, replaying the same draws, must stack the same rows."""
    from oracle import ransac as orz
    import ransac_cases as rc
    gold, vi = _variants()
    assert len(gold["ransac"]) >= 2
    for case in gold["ransac"]:
        pts = vi.terrain_cloud(case["seed"], **{k: tuple(v) if isinstance(v, list) else v for k, v in case["cloud"].items()})
        assert len(pts) == case["n"] and vi.digest(pts) == case["points_sha256"]
        tri = rc.golden_run_triples(pts, case)
        ng, g, _ = orz.remove_ground_tiled_ransac(pts, case["tile_size"], case["distance_threshold"], case["max_iterations"],
                                                  triples={t: tri[t] for t in range(len(tri))})
        assert len(g) == case["ground"]["rows"] and vi.digest(g) == case["ground"]["sha256"]
        assert len(ng) == case["non_ground"]["rows"] and vi.digest(ng) == case["non_ground"]["sha256"]
