"""CPU: host-side logic (LAS container, numpy-percentile restatement, C ABI symbols, generator)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_las_header_roundtrip(tmp_path):
    from pointcloudhookup_b200 import las, synth
    from oracle import las_io
    p = str(tmp_path / "a.las")
    h = synth.write_corridor_las(p, 5000, 2, "flat", 3)
    assert h.point_count == 5000 and h.record_length == 34 and h.point_format == 3 and h.version == (1, 2)
    hdr, rec = las.read_raw(p)
    assert rec.size == 5000 * 34
    o = las_io.read_las(p)
    r = synth.corridor_records(5000, 2, "flat", 3)
    assert np.array_equal(o["X"], r["X"]) and np.array_equal(o["Z"], r["Z"])
    assert np.allclose(hdr.mins, [o["X"].min() * 0.001 + 437000, o["Y"].min() * 0.001 + 3139000, o["Z"].min() * 0.001])
    nh = las.new_header_like(hdr)
    assert nh.record_length == 34 and nh.offset_to_point_data == 227 and nh.n_vlr == 0
    las.write_raw(str(tmp_path / "b.las"), nh, rec, np.zeros(3), np.ones(3))
    assert las.read_header(str(tmp_path / "b.las")).point_count == 5000


def test_las_errors(tmp_path):
    from pointcloudhookup_b200 import las
    with pytest.raises(FileNotFoundError):
        las.read_raw(str(tmp_path / "nope.las"))
    bad = tmp_path / "bad.las"
    bad.write_bytes(b"NOPE" + b"\0" * 400)
    with pytest.raises(las.LasError):
        las.read_raw(str(bad))
    from pointcloudhookup_b200 import synth
    p = str(tmp_path / "c.las")
    synth.write_corridor_las(p, 100, 1, "flat", 1)
    raw = bytearray(open(p, "rb").read())
    raw[104] |= 0x80  # LAZ flag
    (tmp_path / "laz.las").write_bytes(bytes(raw))
    with pytest.raises(las.LasError):
        las.read_raw(str(tmp_path / "laz.las"))
    (tmp_path / "trunc.las").write_bytes(open(p, "rb").read()[:-10])
    with pytest.raises(las.LasError):
        las.read_raw(str(tmp_path / "trunc.las"))


def test_las14_header(tmp_path):
    from pointcloudhookup_b200 import las
    h = las.LasHeader(version=(1, 4), point_format=6, record_length=30, header_size=375, offset_to_point_data=375,
                      point_count=7, scales=np.array([0.01] * 3), offsets=np.zeros(3))
    p = tmp_path / "v14.las"
    p.write_bytes(las.header_bytes(h) + b"\0" * (7 * 30))
    g = las.read_header(str(p))
    assert g.point_count == 7 and g.version == (1, 4) and g.point_format == 6 and g.record_length == 30


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 100, 101, 1000, 4097, 99999, 2**24 + 3, 40_000_001])
@pytest.mark.parametrize("q", [25, 10, 20, 50, 0, 100, 33.3])
def test_percentile_restatement_matches_numpy(n, q):
    """percentile_ranks_f32 + percentile_lerp_f32 == np.percentile for float32 input, including the
    n > 2^24 regime where numpy's float32 virtual index loses integer precision."""
    from pointcloudhookup_b200 import towers as tw
    if n > 10**6:
        z = np.arange(n, dtype=np.float32)   # sorted, distinct enough; value == index up to 2^24
        z *= np.float32(0.37)
    else:
        z = np.random.default_rng(n).normal(0, 10, n).astype(np.float32)
    r0, r1, gamma = tw.percentile_ranks_f32(n, q)
    s = np.sort(z)
    got = tw.percentile_lerp_f32(s[r0], s[r1], gamma)
    exp = np.percentile(z, q)
    assert got.dtype == np.float32 and got == exp
    assert (got + 3.0).dtype == np.float32


def test_c_abi_exports_every_declared_symbol():
    """libpch_b200.so loads and exports every function include/pch_b200.h declares; the ctypes table
    binds exactly that set (no compute call is made — there is no GPU here)."""
    from pointcloudhookup_b200 import _native
    header = open(os.path.join(ROOT, "include", "pch_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pch_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared but not exported"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert _native.lib().pch_version() >= 100
    assert ctypes.sizeof(_native.VoxelPlan) == 32 and ctypes.sizeof(_native.ClusterStats) == 56
    assert ctypes.sizeof(_native.GeoidGrid) == 48 and ctypes.sizeof(_native.TmParams) == 96


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pointcloudhookup_b200 import _native, device as dv
    with pytest.raises(_native.NativeError):
        dv.upload_records(np.zeros(34, np.uint8), 1, 34, [1, 1, 1], [0, 0, 0])
    from pointcloudhookup_b200.ui import import_PC
    with pytest.raises(Exception):
        import_PC.process_chunk(np.zeros((4, 3)), 0.1)
    from pointcloudhookup_b200 import ground_ransac, tiles
    with pytest.raises(_native.NativeError):
        ground_ransac.remove_ground_tiled_ransac(np.zeros((100, 3)))
    with pytest.raises(_native.NativeError):
        tiles.DeviceClusterer()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pointcloudhookup_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_synth_tower_truth_and_order():
    from pointcloudhookup_b200 import synth
    r = synth.corridor_records(60000, 3, "hilly", 9)
    gt = synth.tower_ground_truth(3, "hilly", 9)
    assert gt.shape == (3, 5) and np.all((gt[:, 3] >= 25) & (gt[:, 3] <= 45))
    tower = r[r["classification"] == 15]
    assert 0.04 * 60000 < tower.size < 0.06 * 60000
    # flight order: along-axis coordinate is non-decreasing inside every block
    az = np.radians(synth.AZIMUTH_DEG)
    e = r["X"] * 0.001 + 437000 - synth.ORIGIN_EN[0]
    n = r["Y"] * 0.001 + 3139000 - synth.ORIGIN_EN[1]
    s = e * np.sin(az) + n * np.cos(az)
    assert np.all(np.diff(s[:20000]) > -1e-2)


def test_dropin_surface_matches_reference_names():
    import importlib
    import inspect
    import pointcloudhookup_b200 as pkg
    want = {
        "ui.import_PC": ["process_chunk", "run_voxel_downsampling"],
        "ui.Sampling": ["process_chunk", "voxel_downsample_open3d"],
        "ui.extract": ["create_bbox_using_kuangxuan_method", "create_bbox_lineset_from_bounds",
                       "extract_and_visualize_towers_kuangxuan", "create_enhanced_tower_boxes_kuangxuan",
                       "BBOX_PRESETS", "get_bbox_preset", "visualize_towers_with_point_cloud_kuangxuan",
                       "extract_and_visualize_towers_original", "extract_and_visualize_towers"],
        "ui.compress": ["GIMUtils", "utils", "GIMExtractor"],
        "utils.tower_extraction": ["extract_towers", "_save_tower_las", "create_obb_geometries", "extract_towers_optimized"],
        "utils.elevation_converter": ["ElevationConverter", "convert_elevation"],
        "utils.table_match_gim": ["haversine", "convert_pointcloud_ellipsoid_to_orthometric", "match_towers"],
        "crs": ["ellipsoid_to_orthometric_egm96", "cgcs2000_gk114_to_wgs84"],
    }
    for mod, names in want.items():
        m = importlib.import_module(f"{pkg.__name__}.{mod}")
        for n in names:
            assert hasattr(m, n), (mod, n)
    from pointcloudhookup_b200.utils.tower_extraction import extract_towers
    sig = inspect.signature(extract_towers)
    assert list(sig.parameters)[:10] == ["input_las_path", "progress_callback", "log_callback", "eps", "min_points",
                                          "aspect_ratio_threshold", "min_height", "max_width", "min_width",
                                          "duplicate_threshold"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["eps"], d["min_points"], d["aspect_ratio_threshold"], d["min_height"], d["max_width"], d["min_width"],
            d["duplicate_threshold"]) == (8.0, 80, 0.8, 15.0, 50.0, 8, 30.0)
    from pointcloudhookup_b200.ui.import_PC import run_voxel_downsampling
    d = {k: v.default for k, v in inspect.signature(run_voxel_downsampling).parameters.items()}
    assert d["voxel_size"] == 0.1 and d["chunk_size"] == 1000000
    with pytest.raises(FileNotFoundError):
        run_voxel_downsampling("/nonexistent/in.las", "/tmp/x/out.las")


def test_extract_box_geometry():
    from pointcloudhookup_b200.ui import extract
    lo, hi = extract.create_bbox_using_kuangxuan_method([10.0, 20.0, 30.0], 20.0, 17.0)
    # test/kuangxuan.py:69-71: x in [cx - w, cx + w/0.6 (~1.67w)], y in [cy - w/2, cy + w], z in [cz - h, cz + 2h]
    assert np.allclose(lo, [-10.0, 10.0, 13.0]) and np.allclose(hi, [10 + 33.4, 40.0, 64.0])
    pts, col = extract.create_bbox_lineset_from_bounds(lo, hi)
    assert pts.shape == (24, 3) and col == (1.0, 0.0, 0.0)
    assert extract.get_bbox_preset("nope") == extract.get_bbox_preset("kuangxuan_original")
    geo = extract.create_enhanced_tower_boxes_kuangxuan([{"center": np.array([0.0, 0, 0]), "extent": [10, 12, 30]}])
    assert [g[0].shape for g in geo] == [(24, 3), (24, 3), (2, 3)]
    from pointcloudhookup_b200.utils.tower_extraction import create_obb_geometries
    g = create_obb_geometries([{"center": np.zeros(3), "extent": np.array([2.0, 4, 6]), "rotation": np.eye(3)}])
    assert g[0].points.shape == (8, 3) and g[0].lines.shape == (12, 2)
    assert np.allclose(np.ptp(g[0].points, axis=0), [2, 4, 6])


def test_host_pack_xyz_gathers_the_first_twelve_bytes_of_every_record():
    """pch_host_pack_xyz is host code (no GPU needed): every record length / thread count / ragged tail."""
    from pointcloudhookup_b200 import _native
    lib = _native.lib()
    for rec_len in (12, 20, 26, 28, 34, 37, 67):
        for n in (0, 1, 3, 5, 1000, 70001, 262147):
            rng = np.random.default_rng(n + rec_len)
            src = rng.integers(0, 256, size=n * rec_len + 16, dtype=np.uint8)
            dst = np.zeros(n * 12 + 32, dtype=np.uint8)
            off = (-dst.ctypes.data) % 16
            for threads in (1, 4):
                dst[:] = 0
                assert lib.pch_host_pack_xyz(src.ctypes.data, n, rec_len, dst.ctypes.data + off, threads) == 0
                want = src[: n * rec_len].reshape(n, rec_len)[:, :12].reshape(-1)
                assert np.array_equal(dst[off: off + n * 12], want), (rec_len, n, threads)
                assert not dst[off + n * 12:].any()
    assert lib.pch_host_pack_xyz(None, 5, 34, None, 1) != 0 and b"null" in lib.pch_last_error()


def test_staging_pool_hands_out_each_buffer_once(monkeypatch):
    """The pinned staging pool (device.acquire_staging / release_staging, shared by pipeline and upload_records_xyz) is host logic: buffers are matched by
    identity (tensors compare elementwise), the smallest fitting one is reused, and nothing is handed out twice."""
    import torch
    from pointcloudhookup_b200 import device as dvm, pipeline
    made = []

    def fake(nbytes):
        t = torch.zeros(nbytes, dtype=torch.uint8)
        made.append(t)
        return t
    monkeypatch.setattr(dvm, "_alloc_pinned", fake)
    monkeypatch.setattr(dvm, "_POOL", [])
    a = pipeline._acquire_staging(3_600_000)
    b = pipeline._acquire_staging(3_720_000)
    assert a is not b and len(made) == 2
    pipeline._release_staging(a)
    pipeline._release_staging(b)
    pipeline._release_staging(b)                       # double release is ignored
    assert len(dvm._POOL) == 2
    c = pipeline._acquire_staging(3_600_001)           # only b fits
    assert c is b and len(made) == 2
    d = pipeline._acquire_staging(100)                 # smallest fitting = a
    assert d is a and not dvm._POOL
    e = pipeline._acquire_staging(100)
    assert e is not a and e is not b and len(made) == 3 and e.numel() == 100
    for t in (a, b, e, pipeline._acquire_staging(5), pipeline._acquire_staging(6), pipeline._acquire_staging(7)):
        pipeline._release_staging(t)
    assert len(dvm._POOL) == 4                    # the pool keeps the four largest
    pipeline._release_staging(None)


def test_merge_adjacent_clusters_matches_the_reference_variant():
    """towers.merge_adjacent_clusters (on per-cluster stats) against the literal restatement of test/tttt.py:93-175
    (on points + labels, real sklearn KDTree): same partition, same new-label order, merged stats = stats of the
    relabelled points."""
    from oracle import towers as ot
    from pointcloudhookup_b200 import towers as tw, device as dv
    rng = np.random.default_rng(12)
    for trial in range(6):
        K = int(rng.integers(1, 60))
        centres = rng.uniform(0, 60, (K, 3)) * np.array([1.0, 1.0, 0.2])
        sizes = rng.integers(3, 200, K)
        pts, labs = [], []
        for k in range(K):
            pts.append((centres[k] + rng.normal(0, 0.4, (sizes[k], 3))).astype(np.float32))
            labs.append(np.full(sizes[k], k, dtype=np.int32))
        pts.append(rng.uniform(0, 60, (50, 3)).astype(np.float32))
        labs.append(np.full(50, -1, dtype=np.int32))
        pts, labs = np.concatenate(pts), np.concatenate(labs)
        perm = rng.permutation(len(labs))
        pts, labs = pts[perm], labs[perm]
        stats = np.zeros(K, dtype=dv.STATS_DTYPE)
        for k in range(K):
            cp = pts[labs == k]
            stats[k] = (len(cp), cp.min(0), cp.max(0), cp.astype(np.float64).sum(0))
        thr = 6.0
        want = ot.merge_adjacent_clusters(pts, labs, thr)
        comp, merged = tw.merge_adjacent_clusters(stats, K, thr)
        base = labs.max() + 1
        got = np.where(labs >= 0, base + comp[np.maximum(labs, 0)], -1)
        assert np.array_equal(got, want), trial
        for c in range(len(merged)):
            cp = pts[want == base + c]
            assert merged["count"][c] == len(cp)
            assert np.array_equal(merged["min"][c], cp.min(0)) and np.array_equal(merged["max"][c], cp.max(0))
            assert np.allclose(merged["sum"][c], cp.astype(np.float64).sum(0), rtol=0, atol=1e-6)
    comp, merged = tw.merge_adjacent_clusters(np.zeros(0, dtype=dv.STATS_DTYPE), 0, 6.0)
    assert len(comp) == 0 and len(merged) == 0


def test_select_towers_with_merge_threshold_uses_merged_boxes():
    """select_towers(merge_threshold=...) in AABB mode is pure host logic on the cluster stats: two halves of a tower
    that fail the size filter alone pass once merged, and carry the reference's max(label)+1 numbering."""
    from pointcloudhookup_b200 import towers as tw, device as dv
    stats = np.zeros(3, dtype=dv.STATS_DTYPE)
    # lower and upper half of one tower (each 12 x 12 x 10 m: too short alone), and a far-away bush
    stats[0] = (500, (0, 0, 0), (12, 12, 10), (500 * 6.0, 500 * 6.0, 500 * 5.0))
    stats[1] = (400, (0, 0, 10), (12, 12, 20), (400 * 6.0, 400 * 6.0, 400 * 9.0))     # centre (6,6,9): 4 m from (6,6,5)
    stats[2] = (300, (200, 0, 0), (203, 3, 2), (300 * 201.5, 300 * 1.5, 300 * 1.0))
    st = tw.TowerStages(None, np.array([1000.0, 2000.0, 50.0], dtype=np.float32), np.float32(0), 3.0, None, None, 3, stats)
    assert tw.select_towers(st, box="aabb", want_points=False) == []
    got = tw.select_towers(st, box="aabb", want_points=False, merge_threshold=6.0)
    assert [t["label"] for t in got] == [3]                       # first merged component -> max(label)+1
    assert np.allclose(got[0]["extent"], [12, 12, 20]) and np.allclose(got[0]["center"], [1006, 2006, 60])
    assert tw.select_towers(st, box="aabb", want_points=False, merge_threshold=1.0) == []


def test_pack_mode_resolution(monkeypatch):
    from pointcloudhookup_b200 import pipeline
    assert pipeline.resolve_pack("xyz") == "xyz" and pipeline.resolve_pack("none", 64) == "none"
    assert pipeline.resolve_pack("auto", 16) == "xyz" and pipeline.resolve_pack("auto", 4) == "none"
    monkeypatch.setattr(pipeline, "host_threads", lambda: 2)
    assert pipeline.resolve_pack("auto") == "none"
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    from pointcloudhookup_b200 import device as dv
    assert dv.host_threads() >= 1


def test_slice_plan_of_the_host_entries():
    """pipeline.slice_plan is pure host logic: whole-chunk slices, 16-byte aligned starts in both layouts, the
    raw-slice pattern only for pinned sources."""
    from pointcloudhookup_b200 import pipeline
    cs, b, l = pipeline.slice_plan(1_200_001, 34, 500_000, 1, "xyz", 0, False)
    assert cs == 500_000 and b == [(0, 500_000), (500_000, 1_000_000), (1_000_000, 1_200_001)] and l == [12, 12, 12]
    cs, b, l = pipeline.slice_plan(1_200_001, 34, 500_000, 2, "none", 3, True)
    assert b == [(0, 1_000_000), (1_000_000, 1_200_001)] and l == [34, 34]
    cs, b, l = pipeline.slice_plan(100, 34, 8, 1, "xyz", 3, True)            # every third slice as whole records
    assert len(b) == 13 and l[:6] == [12, 12, 34, 12, 12, 34] and all((lo * 12) % 16 == 0 and (lo * 34) % 16 == 0 for lo, _ in b)
    assert pipeline.slice_plan(100, 34, 8, 1, "xyz", 3, False)[2] == [12] * 13                        # pageable source: no raw slices
    cs, b, l = pipeline.slice_plan(100, 34, 6, 1, "xyz", 0, True)            # 6*34 % 16 != 0 -> one slice
    assert cs == 6 and b == [(0, 100)] and l == [12]
    cs, b, l = pipeline.slice_plan(4, 36, 50_000, 2, "xyz", 0, False)        # tile smaller than a chunk
    assert cs == 4 and b == [(0, 4)]
    assert pipeline.slice_plan(0, 34, 500_000, 10, "none", 0, True)[1] == []
    with pytest.raises(ValueError):
        pipeline.slice_plan(10, 34, 5, 1, "zip", 0, True)


def test_grid_shape_caps_the_cell_table():
    """device.grid_shape: the grid-min cell table must stay in proportion to the cloud (one far outlier would
    otherwise ask for gigabytes) and inside the int32 cell index of the C ABI."""
    import pytest
    from pointcloudhookup_b200 import device as dvm
    assert dvm.grid_shape([0.0, 0.0], [299.9, 59.9], 2.0, 200000) == (150, 30)
    assert dvm.grid_shape([-5.0, 1.0], [-5.0, 1.0], 2.0, 1) == (1, 1)
    with pytest.raises(ValueError, match="out of proportion"):
        dvm.grid_shape([0.0, 0.0], [1.0e7, 1.0e7], 2.0, 1000)
    with pytest.raises(ValueError):
        dvm.grid_shape([0.0, 0.0], [np.inf, 1.0], 2.0, 1000)
    with pytest.raises(ValueError):
        dvm.grid_shape([0.0, 0.0], [1.0, 1.0], 0.0, 1000)
    # a long corridor is fine: 17.5 km x 5.5 km bounding box at 2 m cells for 78 M points
    assert dvm.grid_shape([0.0, 0.0], [5500.0, 17500.0], 2.0, 78_000_000) == (2751, 8751)


def test_integration_stub_matches_the_abi():
    """INTEGRATION.md §2 is executable documentation: every lib.pch_* call in the stub must pass as many arguments as
    the ABI takes (round 1 shipped a call that was one short), and every argtypes list must equal _native's."""
    import ast
    import ctypes as C
    import re
    from pointcloudhookup_b200 import _native
    text = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()
    sec = text[text.index("## 2. Binding the C ABI directly"):text.index("## 3. Entry point")]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    tree = ast.parse(code)
    calls, argtypes = {}, {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                and node.func.value.id == "lib" and node.func.attr.startswith("pch_"):
            calls.setdefault(node.func.attr, []).append(len(node.args))
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Attribute) and node.targets[0].attr == "argtypes":
            argtypes[node.targets[0].value.attr] = len(node.value.elts)
    assert "pch_voxel_downsample_las" in calls and "pch_las_chunk_minmax" in calls
    for name, counts in calls.items():
        assert name in _native.SIGNATURES, name
        for c in counts:
            assert c == len(_native.SIGNATURES[name][1]), f"{name}: the stub passes {c} arguments, the ABI takes {len(_native.SIGNATURES[name][1])}"
    for name, c in argtypes.items():
        assert c == len(_native.SIGNATURES[name][1]), name
    # every function the stub CALLS with pointers has argtypes set (else 64-bit pointers are truncated to c_int)
    assert {"pch_voxel_downsample_las", "pch_las_chunk_minmax"} <= set(argtypes)
    # the argtypes section runs against the real library
    head = code[:code.index("def check")].replace('C.CDLL("pointcloudhookup_b200/libpch_b200.so")', f'C.CDLL({_native.LIB_PATH!r})')
    head = head.replace("import ctypes as C, torch, numpy as np", "import ctypes as C")
    ns = {}
    exec(head, ns)
    sig = _native.SIGNATURES["pch_voxel_downsample_las"][1]
    assert [C.sizeof(t) for t in ns["lib"].pch_voxel_downsample_las.argtypes] == [C.sizeof(t) for t in sig]


def test_arange_edges3_reproduces_numpy_arange():
    """pch_ransac_tile_words rebuilds the tile edges as e[0] + i*(e[1]-e[0]) with e[1] taken as given: that is
    np.arange's fill rule, bit for bit, also where (start + step) - start != step."""
    from pointcloudhookup_b200 import ground_ransac as gr
    rng = np.random.default_rng(4)
    for _ in range(200):
        lo = float(rng.choice([rng.uniform(-1, 1), rng.uniform(4e5, 6e5), rng.uniform(-3e6, 3e6)]))
        step = float(rng.choice([10.0, 20.0, 0.1, 7.3]))
        hi = lo + float(rng.uniform(0, 40)) * step
        e = np.arange(lo, hi, step)
        n, e3 = gr.arange_edges3(lo, hi, step)
        assert n == len(e)
        if n < 2:
            assert e3 is None
            continue
        assert e3[0] == e[0] and e3[1] == e[1]
        rebuilt = [e3[0] if i == 0 else e3[1] if i == 1 else e3[0] + i * e3[2] for i in range(n)]
        assert np.array_equal(np.array(rebuilt), e)
