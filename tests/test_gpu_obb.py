"""GPU: minimum-volume oriented boxes on the device (pch_obb_batch: gift-wrapping hull + every face normal) against
the independent oracle (oracle.obb.min_volume_box_all_faces: Qhull in 3-D and 2-D + numpy).  Tolerance 1e-4 m on
extents and centre (north_star's bound for float outputs); axes compared after the shared sign convention.
Reference call site: utils/tower_extraction.py:137-151."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clusters(rng):
    out = []
    for trial in range(10):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        n = int(rng.integers(300, 30000))
        p = (rng.uniform(-1, 1, (n, 3)) * rng.uniform(2, 25, 3)) @ q.T + rng.uniform(-80, 80, 3)
        if trial % 3 == 1:
            p = np.concatenate([p, rng.normal(0, 2, (n // 2, 3)) + p.mean(0)])          # a dense core inside
        if trial % 3 == 2:
            p = np.round(p, 1)                                                          # lattice: many coplanar / collinear points
        out.append(p.astype(np.float32))
    return out


def test_obb_batch_matches_oracle(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv, synth, towers as tw
    from oracle import obb
    rng = np.random.default_rng(17)
    clusters = _clusters(rng)
    # a synthetic lattice tower (legs, bracing, cross-arms), the shape the reference boxes
    ds, dt, dz = synth._tower_points(np.random.default_rng(3), 20000, 38.0, 11.0)
    clusters.append(np.column_stack([ds * 0.29 + dt * 0.95, -ds * 0.95 + dt * 0.29, dz]).astype(np.float32))
    rows = np.concatenate(clusters)
    off = np.concatenate([[0], np.cumsum([len(c) for c in clusters])])
    res = dv.obb_batch(torch.from_numpy(rows).to(cuda_device), np.stack([off[:-1], off[1:]], axis=1))
    assert res.shape[0] == len(clusters)
    for c, r in zip(clusters, res):
        assert int(r["status"]) == 0, r
        t, ext, vol = obb.min_volume_box_all_faces(c)
        assert abs(r["volume"] - vol) <= 1e-6 * vol, (r["volume"], vol)
        assert np.abs(np.asarray(r["extents"]) - ext).max() < 1e-4
        assert np.abs(np.asarray(r["center"]) - t[:3, 3]).max() < 1e-4
        a0, a1, a2 = tw.canonical_box_axes(r["rotation"][:, 0], r["rotation"][:, 2])
        got = np.column_stack((a0, a1, a2))
        assert np.abs(got - t[:3, :3]).max() < 1e-5
        assert abs(np.linalg.det(got) - 1.0) < 1e-9
        assert 0 < r["n_vertices"] <= r["n_candidates"] <= len(c)
        # every point is inside the device box
        loc = (c.astype(np.float64) - np.asarray(r["center"])) @ np.asarray(r["rotation"])
        assert np.all(np.abs(loc) <= np.asarray(r["extents"]) / 2 + 1e-6)


def test_obb_batch_degenerate_clusters_are_flagged(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(2)
    tiny = rng.normal(size=(3, 3)).astype(np.float32)
    flat = np.column_stack([rng.uniform(0, 10, 500), rng.uniform(0, 5, 500), np.full(500, 2.0)]).astype(np.float32)
    line = np.column_stack([np.linspace(0, 10, 200), np.linspace(0, 5, 200), np.linspace(1, 2, 200)]).astype(np.float32)
    same = np.tile(np.array([[1.0, 2.0, 3.0]], dtype=np.float32), (50, 1))
    ok = (rng.uniform(-1, 1, (800, 3)) * [3, 4, 5]).astype(np.float32)
    cl = [tiny, flat, line, same, ok]
    off = np.concatenate([[0], np.cumsum([len(c) for c in cl])])
    res = dv.obb_batch(torch.from_numpy(np.concatenate(cl)).to(cuda_device), np.stack([off[:-1], off[1:]], axis=1))
    assert [int(s) for s in res["status"][:4]] == [1, 2, 2, 1]
    assert int(res["status"][4]) == 0 and res["volume"][4] > 0


def test_select_towers_obb_on_device_matches_oracle(cuda_device):
    """The production call (box="obb", the reference default): boxes from the device for every candidate cluster in
    one launch; same towers as the oracle's exhaustive hull-face search."""
    import torch
    from pointcloudhookup_b200 import device as dv, synth, towers as tw
    from oracle import las_io, towers as ot, voxel as ov
    n = 400_000
    rec = synth.corridor_records(n, 3, "flat", 31, (0.86, 0.085, 0.005, 0.05))
    las = {"scales": synth.SCALES, "offsets": synth.OFFSETS, "X": rec["X"].copy(), "Y": rec["Y"].copy(), "Z": rec["Z"].copy(), "n": n}
    ref, _ = ov.downsample_las_arrays(las, 0.1, 200_000)
    q = [las_io.quantise(ref[:, i], synth.SCALES[i], synth.OFFSETS[i]) for i in range(3)]
    las2 = dict(las, X=q[0], Y=q[1], Z=q[2], n=len(q[0]))
    exp = ot.extract_towers_arrays(las2, box="obb")
    dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
    raw = dv.voxel_downsample(dl, 0.1, 200_000, want=("f32",)).f32
    stages = tw.run_stages(raw)
    got = tw.select_towers(stages, box="obb")
    assert len(exp) >= 2 and [t["label"] for t in got] == [t["label"] for t in exp]
    for a, b in zip(got, exp):
        assert np.abs(a["center"] - b["center"]).max() < 1e-4 and np.abs(a["extent"] - b["extent"]).max() < 1e-4
        assert abs(a["north_angle"] - b["north_angle"]) < 1e-3
        assert np.array_equal(a["points"], b["points"])
