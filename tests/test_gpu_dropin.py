"""GPU: the drop-in modules produce what the UNMODIFIED reference modules produced on the same
inputs (tests/golden/reference_run.json) — files, tower lists, callback sequences, error behaviour."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))
CASES = ["dense_wires", "towers", "hilly"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("dropin")


def make_input(case, workdir):
    from pointcloudhookup_b200 import synth
    c = GOLD[case]["config"]
    path = os.path.join(workdir, f"{case}.las")
    if not os.path.exists(path):
        synth.write_corridor_las(path, c["n"], c["towers"], c["terrain"], c["seed"], tuple(c["fractions"]))
    return path, c


@pytest.mark.parametrize("case", CASES)
def test_run_voxel_downsampling_matches_reference(cuda_device, case, workdir):
    from pointcloudhookup_b200.ui import import_PC, Sampling
    from oracle import las_io
    path, c = make_input(case, workdir)
    out = os.path.join(workdir, "output", f"{case}_ds.las")
    logs, prog = [], []
    import_PC.run_voxel_downsampling(path, out, voxel_size=c["voxel"], chunk_size=c["chunk"],
                                     progress_callback=prog.append, log_callback=logs.append)
    g = GOLD[case]["downsample"]
    ds = las_io.read_las(out)
    assert ds["n"] == g["count"] and prog == g["progress"] and len(logs) == g["n_logs"]
    assert digest(np.stack([ds["X"], ds["Y"], ds["Z"]], 1)) == g["xyz_sha256"]
    assert ds["record_length"] == 34 and ds["point_format"] == 3 and ds["version"] == (1, 2)
    out2 = os.path.join(workdir, "output", f"{case}_ds2.las")
    Sampling.voxel_downsample_open3d(path, out2, c["voxel"], c["chunk"])
    assert open(out, "rb").read()[227:] == open(out2, "rb").read()[227:]
    Sampling.voxel_downsample_open3d(os.path.join(workdir, "missing.las"), out2, c["voxel"], c["chunk"])  # no raise
    with pytest.raises(FileNotFoundError):
        import_PC.run_voxel_downsampling(os.path.join(workdir, "missing.las"), out)
    pts = np.stack(las_io.scaled(las_io.read_las(path), 0, 5000), 1)
    pc = import_PC.process_chunk(pts, c["voxel"])
    assert pc.dtype == np.float64 and digest(pc) == GOLD[case]["process_chunk"]["sha256"]
    assert digest(Sampling.process_chunk(pts, None, c["voxel"])) == GOLD[case]["process_chunk"]["sha256"]


@pytest.mark.parametrize("case", CASES)
def test_extract_towers_matches_reference(cuda_device, case, workdir, monkeypatch):
    from pointcloudhookup_b200.ui import import_PC
    from pointcloudhookup_b200.utils import tower_extraction as te
    from oracle import las_io
    path, c = make_input(case, workdir)
    out = os.path.join(workdir, "output", f"{case}_ds.las")
    if not os.path.exists(out):
        import_PC.run_voxel_downsampling(path, out, voxel_size=c["voxel"], chunk_size=c["chunk"])
    run_dir = os.path.join(workdir, f"run_{case}")
    os.makedirs(run_dir, exist_ok=True)
    monkeypatch.chdir(run_dir)
    logs, prog = [], []
    res = te.extract_towers(out, progress_callback=prog.append, log_callback=logs.append)
    g = GOLD[case]["towers"]
    assert prog == g["progress"]
    assert len(res) == g["count"]
    for a, b in zip(res, g["list"]):
        assert set(a) == {"center", "rotation", "extent", "height", "width", "north_angle", "points"}
        assert np.abs(a["center"] - np.array(b["center"])).max() < 1e-4
        assert np.abs(np.asarray(a["extent"]) - np.array(b["extent"])).max() < 1e-4
        assert np.abs(np.asarray(a["rotation"]) - np.array(b["rotation"])).max() < 1e-6
        assert abs(a["north_angle"] - b["north_angle"]) < 1e-4
        assert a["points"].dtype == np.float32 and digest(a["points"]) == b["points_sha256"]
    assert sorted(os.listdir("output_towers")) == g["tower_files"]
    for f, b in zip(g["tower_files"], g["list"]):
        t = las_io.read_las(os.path.join("output_towers", f))
        assert t["n"] == b["n_points"]
    assert te.extract_towers(os.path.join(workdir, "missing.las"), log_callback=logs.append) == []   # never raises


def test_extract_module_reads_full_cloud(cuda_device, workdir):
    from pointcloudhookup_b200.ui import extract
    from oracle import las_io
    path, _ = make_input("dense_wires", workdir)
    towers = [{"center": np.array([437551.1, 3139667.3, 103.8]), "extent": np.array([35.0, 16.2, 41.4]), "rotation": np.eye(3)}]
    full, geoms = extract.extract_and_visualize_towers(path, towers)
    ref = np.stack(las_io.scaled(las_io.read_las(path)), 1)
    assert full.dtype == np.float64 and np.array_equal(full, ref)
    assert len(geoms) == 1 and geoms[0][0].shape == (24, 3)
    full2, geoms2 = extract.extract_and_visualize_towers(path, towers, use_kuangxuan_method=False)
    assert np.array_equal(full2, ref) and geoms2[0][0].shape == (24, 3)
    with pytest.raises(FileNotFoundError):
        extract.extract_and_visualize_towers(os.path.join(workdir, "nope.las"), towers)


def test_threaded_call_like_the_gui(cuda_device, workdir):
    """pyGUI_towers_test.py:385 runs the hot path on a daemon thread."""
    import threading
    from pointcloudhookup_b200.ui import import_PC
    path, c = make_input("dense_wires", workdir)
    out = os.path.join(workdir, "output", "threaded.las")
    err = []

    def work():
        try:
            import_PC.run_voxel_downsampling(path, out, c["voxel"], c["chunk"])
        except Exception as e:  # pragma: no cover
            err.append(e)
    th = threading.Thread(target=work, daemon=True)
    th.start()
    th.join(120)
    assert not err and os.path.exists(out)


def test_upload_records_xyz_streams_any_host_buffer(cuda_device, tmp_path):
    """Double-buffered gather + H2D of a memmapped LAS file, an ndarray and a tensor: the 12-byte stream decodes
    to exactly what the whole records decode to, for ragged block sizes."""
    import torch
    from pointcloudhookup_b200 import device as dv, las as _las, synth
    n = 123_457
    rec = synth.corridor_records(n, 2, "hilly", 77)
    path = str(tmp_path / "in.las")
    hdr = _las.LasHeader(point_format=3, record_length=34, scales=np.array(synth.SCALES), offsets=np.array(synth.OFFSETS))
    _las.write_raw(path, hdr, rec.view(np.uint8).reshape(-1))
    h2, mm = _las.read_raw(path)
    full = dv.upload_records(rec.view(np.uint8).reshape(-1), n, 34, synth.SCALES, synth.OFFSETS)
    want = dv.decode_xyz(full, torch.float64)
    for src, block, threads in ((mm, 1 << 22, 0), (mm, 10_000, 3), (rec.view(np.uint8).reshape(-1).copy(), 40_001, 1),
                                (torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()), 4, 2)):
        dl = dv.upload_records_xyz(src, n, 34, h2.scales, h2.offsets, block_points=block, threads=threads)
        assert dl.rec_len == 12 and torch.equal(dv.decode_xyz(dl, torch.float64), want)
    e = dv.upload_records_xyz(np.zeros(0, np.uint8), 0, 34, synth.SCALES, synth.OFFSETS)
    assert e.n == 0 and dv.decode_xyz(e, torch.float64).shape == (0, 3)
    with pytest.raises(ValueError):
        dv.upload_records_xyz(np.zeros(33, np.uint8), 1, 34, synth.SCALES, synth.OFFSETS)
