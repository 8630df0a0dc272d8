"""GPU edge cases: empty and tiny inputs, other LAS versions / record formats, extra bytes, negative
coordinates and scales, degenerate voxel sizes, error paths."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_las(path, X, Y, Z, version=(1, 2), pfmt=3, rec_len=None, scales=(0.01, 0.01, 0.01), offsets=(0.0, 0.0, 0.0),
               pad_header=0):
    from pointcloudhookup_b200 import las
    rec_len = rec_len or las.PDRF_LENGTH[pfmt]
    hsize = 375 if version >= (1, 4) else 227
    h = las.LasHeader(version=version, point_format=pfmt, record_length=rec_len, header_size=hsize,
                      offset_to_point_data=hsize + pad_header, point_count=len(X), scales=np.array(scales, float),
                      offsets=np.array(offsets, float))
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, size=(len(X), rec_len), dtype=np.uint8)
    raw[:, :12] = np.stack([X, Y, Z], 1).astype("<i4").view(np.uint8).reshape(len(X), 12)
    las.write_raw(path, h, raw.reshape(-1))
    return h


@pytest.mark.parametrize("version,pfmt,rec_len,pad", [((1, 2), 0, None, 0), ((1, 2), 1, None, 3), ((1, 2), 2, None, 0),
                                                       ((1, 4), 6, None, 0), ((1, 4), 7, 40, 1), ((1, 4), 8, None, 0),
                                                       ((1, 2), 3, 35, 0)])
def test_formats_and_extra_bytes(cuda_device, tmp_path, version, pfmt, rec_len, pad):
    from pointcloudhookup_b200.ui import import_PC, extract
    from oracle import las_io, voxel as ov
    rng = np.random.default_rng(pfmt)
    n = 30011
    X, Y, Z = (rng.integers(-200000, 200000, n).astype(np.int32) for _ in range(3))
    src = str(tmp_path / "in.las")
    _write_las(src, X, Y, Z, version, pfmt, rec_len, scales=(0.001, 0.002, 0.01), offsets=(-50.0, 1e6, 3.5), pad_header=pad)
    las = las_io.read_las(src)
    assert np.array_equal(las["X"], X)
    full, _ = extract.extract_and_visualize_towers(src, [])
    assert np.array_equal(full, np.stack(las_io.scaled(las), 1))
    out = str(tmp_path / "o" / "out.las")
    import_PC.run_voxel_downsampling(src, out, voxel_size=0.5, chunk_size=7000)
    ref, _ = ov.downsample_las_arrays(las, 0.5, 7000)
    got = las_io.read_las(out)
    q = [las_io.quantise(ref[:, i], las["scales"][i], las["offsets"][i]) for i in range(3)]
    assert got["n"] == len(ref) and all(np.array_equal(got[k], q[i]) for i, k in enumerate("XYZ"))
    assert got["version"] == version and got["point_format"] == pfmt
    from pointcloudhookup_b200 import las as plas
    assert got["record_length"] == plas.PDRF_LENGTH[pfmt]          # nominal length, extra bytes dropped (laspy)


def test_empty_and_tiny_inputs(cuda_device, tmp_path):
    import torch
    from pointcloudhookup_b200 import device as dv, towers as tw
    from pointcloudhookup_b200.ui import import_PC
    from pointcloudhookup_b200.utils import tower_extraction as te
    from oracle import las_io
    e = np.zeros(0, np.int32)
    src = str(tmp_path / "empty.las")
    _write_las(src, e, e, e)
    out = str(tmp_path / "o" / "e.las")
    logs = []
    import_PC.run_voxel_downsampling(src, out, log_callback=logs.append)
    assert las_io.read_las(out)["n"] == 0
    assert te.extract_towers(src, log_callback=logs.append, write_outputs=False) == []
    assert import_PC.process_chunk(np.zeros((0, 3)), 0.1).shape == (0, 3)
    one = import_PC.process_chunk(np.array([[1.0, 2.0, 3.0]]), 0.1)
    assert np.array_equal(one, [[1.0, 2.0, 3.0]])
    same = import_PC.process_chunk(np.tile([[5.0, 5.0, 5.0]], (1000, 1)), 0.1)
    assert same.shape == (1, 3) and np.array_equal(same[0], [5.0, 5.0, 5.0])
    # DBSCAN on few points: all noise, no clusters
    pts = torch.rand((50, 3), device=cuda_device) * 100
    r = dv.dbscan_chunked(pts.contiguous(), 8.0, 80, 50000)
    assert r.n_clusters == 0 and bool((r.labels == -1).all())
    # fewer than 1000 survivors -> the reference's fallback threshold (base + 1.0)
    z = np.concatenate([np.zeros(3000), np.full(500, 2.0), np.full(200, 10.0)]).astype(np.float32)
    raw = np.stack([np.arange(z.size, dtype=np.float32), np.zeros_like(z), z], 1)
    filt, cen, base, used, _ = tw.ground_filter_percentile(torch.from_numpy(raw).to(cuda_device))
    assert used == 1.0 and filt.shape[0] == 700


def test_process_chunk_matches_oracle_on_arbitrary_floats(cuda_device):
    from pointcloudhookup_b200.ui import import_PC
    from oracle import voxel as ov
    rng = np.random.default_rng(8)
    for n, v in ((10, 0.3), (5000, 0.05), (200000, 1.7), (100000, 0.02)):
        pts = rng.normal(0, 30, (n, 3)) + np.array([-1e5, 2e6, 10.0])
        pts[: n // 4] = np.round(pts[: n // 4], 1)        # exact voxel-boundary / duplicate values
        got = import_PC.process_chunk(pts, v)
        assert np.array_equal(got, ov.voxel_down_sample(pts, v)), (n, v)
    # index range wider than one 64-bit sort word (18+18+18 voxel bits + 17 index bits): wide-key path
    got = import_PC.process_chunk(pts, 1e-3)
    assert np.array_equal(got, ov.voxel_down_sample(pts, 1e-3))
    got32 = import_PC.process_chunk(pts.astype(np.float32), 0.5)   # astype(float64) of float32 input
    assert np.array_equal(got32, ov.voxel_down_sample(pts.astype(np.float32).astype(np.float64), 0.5))


def test_wide_keys_through_las_path(cuda_device):
    """A LAS chunk whose extent/voxel ratio needs more than 64 key bits takes the wide-key fallback and still
    reproduces the oracle (means, re-quantised lattice, float32 read-back, per-chunk counts)."""
    from pointcloudhookup_b200 import device as dv
    from oracle import las_io, voxel as ov
    from conftest import make_las_dict
    rng = np.random.default_rng(12)
    n = 60000
    rec = np.zeros(n, dtype=np.dtype([("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("pad", "u1", 22)]))
    for k in "XYZ":
        rec[k] = rng.integers(-2**30, 2**30, n)
    rec["X"][: n // 3] = rec["X"][0] + rng.integers(0, 3, n // 3)     # some points share voxels
    rec["Y"][: n // 3] = rec["Y"][0]
    rec["Z"][: n // 3] = rec["Z"][0]
    sc, of = np.array([0.001, 0.001, 0.001]), np.array([0.0, 0.0, 0.0])
    dl = dv.upload_records(rec.view(np.uint8), n, 34, sc, of)
    res = dv.voxel_downsample(dl, 0.01, 25000, want=("mean", "lattice", "f32"))   # 2^31 mm / 10 mm -> 28 bits per axis
    las = make_las_dict(rec, sc, of)
    ref, cnt = ov.downsample_las_arrays(las, 0.01, 25000)
    assert res.count == len(ref) and np.array_equal(res.chunk_counts.cpu().numpy(), cnt)
    assert np.array_equal(res.mean.cpu().numpy(), ref)
    lat = np.stack([las_io.quantise(ref[:, i], sc[i], of[i]) for i in range(3)], 1)
    assert np.array_equal(res.lattice.cpu().numpy(), lat)
    assert np.array_equal(res.f32.cpu().numpy(), (lat.astype(np.float64) * sc + of).astype(np.float32))


def test_negative_scale_and_offsets(cuda_device):
    from pointcloudhookup_b200 import device as dv
    from oracle import voxel as ov
    from conftest import make_las_dict
    rng = np.random.default_rng(2)
    n = 40000
    rec = np.zeros(n, dtype=np.dtype([("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("pad", "u1", 22)]))
    for k in "XYZ":
        rec[k] = rng.integers(-300000, 300000, n)
    sc, of = np.array([-0.001, 0.001, 0.0005]), np.array([1234.5, -999.25, 0.0])
    dl = dv.upload_records(rec.view(np.uint8), n, 34, sc, of)
    res = dv.voxel_downsample(dl, 0.25, 9000, want=("mean",))
    ref, cnt = ov.downsample_las_arrays(make_las_dict(rec, sc, of), 0.25, 9000)
    assert np.array_equal(res.mean.cpu().numpy(), ref) and np.array_equal(res.chunk_counts.cpu().numpy(), cnt)


def test_errors(cuda_device, tmp_path):
    import torch
    from pointcloudhookup_b200 import _native, device as dv
    from pointcloudhookup_b200.ui import import_PC
    with pytest.raises(ValueError):
        import_PC.process_chunk(np.zeros((4, 3)), 0.0)             # open3d: voxel_size <= 0
    with pytest.raises(ValueError):
        import_PC.process_chunk(np.zeros((4, 2)), 0.1)
    # open3d's own limit: the voxel index must fit an int
    pts = np.array([[0.0, 0.0, 0.0], [1e9, 1e9, 1e9]])
    with pytest.raises(ValueError):
        import_PC.process_chunk(pts, 1e-4)
    bad = torch.zeros(100, dtype=torch.uint8, device=cuda_device)[1:]   # misaligned record buffer
    dl = dv.DeviceLas(bad, 2, 34, np.ones(3), np.zeros(3))
    with pytest.raises(_native.NativeError):
        dv.decode_xyz(dl)
    src = str(tmp_path / "x.laz.las")
    _write_las(src, np.zeros(4, np.int32), np.zeros(4, np.int32), np.zeros(4, np.int32))
    raw = bytearray(open(src, "rb").read())
    raw[104] |= 0x80
    open(src, "wb").write(bytes(raw))
    with pytest.raises(Exception):
        import_PC.run_voxel_downsampling(src, str(tmp_path / "o" / "y.las"))


def test_compaction_variants_agree_with_numpy(cuda_device):
    """The three compaction kernels (generic, staged xyz, flags-only) on ragged sizes and every keep source."""
    import torch
    from pointcloudhookup_b200 import _native, device as dv
    lib = _native.lib()
    rng = np.random.default_rng(17)
    st = lambda: torch.cuda.current_stream().cuda_stream
    for m in (1, 15, 16, 17, 2047, 2048, 2049, 8191, 8192, 8193, 100_003):
        xyz = rng.normal(0, 50, (m, 3)).astype(np.float32)
        cen = np.array([1.5, -2.25, 3.0], dtype=np.float32)
        d = torch.from_numpy(xyz).to(cuda_device)
        c = torch.from_numpy(cen).to(cuda_device)
        thr = 4.0
        # (a) keep flag derived from the cloud, (b) from a z column, (c) from a byte mask with arbitrary non-zero values
        keep = (xyz[:, 2] - cen[2]) > np.float32(thr)
        got, g, src, mask = dv.compact_points(d, None, thr, c, want_src=True, want_mask=True)
        assert g == keep.sum() and np.array_equal(got.cpu().numpy(), (xyz - cen)[keep])
        assert np.array_equal(src.cpu().numpy(), np.nonzero(keep)[0]) and np.array_equal(mask.cpu().numpy().astype(bool), keep)
        zs = torch.from_numpy((xyz[:, 2] - cen[2]).astype(np.float32)).to(cuda_device)
        got2, g2, _, _ = dv.compact_points(d, zs, thr, c)
        assert g2 == g and torch.equal(got2, got)
        bm = (rng.random(m) < 0.3) * rng.integers(1, 256, m)
        bmask = torch.from_numpy(bm.astype(np.uint8)).to(cuda_device)
        got3, g3, src3, _ = dv.compact_points(d, None, 0.0, None, keep_mask=bmask, want_src=True)
        assert g3 == (bm != 0).sum() and np.array_equal(got3.cpu().numpy(), xyz[bm != 0])
        assert np.array_equal(src3.cpu().numpy(), np.nonzero(bm)[0])
        # flags-only entry (no rows out): the DBSCAN head list path
        for dens in (0.3, 0.0005):
            bm = (rng.random(m) < dens) * rng.integers(1, 256, m)
            bmask = torch.from_numpy(bm.astype(np.uint8)).to(cuda_device)
            out = torch.full((m,), -1, dtype=torch.int32, device=cuda_device)
            cnt = torch.zeros(1, dtype=torch.int64, device=cuda_device)
            wsb = lib.pch_compact_workspace_bytes(m)
            ws = torch.empty(wsb, dtype=torch.uint8, device=cuda_device)
            rc = lib.pch_compact_points(d.data_ptr(), None, bmask.data_ptr(), m, None, 0.0, None, out.data_ptr(), None,
                                        cnt.data_ptr(), ws.data_ptr(), wsb, st())
            assert rc == 0
            k = int(cnt.item())
            assert k == (bm != 0).sum() and np.array_equal(out[:k].cpu().numpy(), np.nonzero(bm)[0])
            assert bool((out[k:] == -1).all())
