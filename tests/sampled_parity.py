"""Oracle parity on SAMPLED units of a full-size run (BASELINE.json configs[1]/[2] sizes).

The reference's units are independent: voxel chunks of `chunk_size` consecutive points
(ui/import_PC.py:45-58) and DBSCAN chunks of 50 000 consecutive filtered points
(utils/tower_extraction.py:96-116).  So a handful of randomly drawn chunks of a 20 M / 100 M-point run can
be checked against the oracle (numpy / real scikit-learn) in seconds, bit for bit, while centroid,
percentile base and the keep mask are cheap enough to check against numpy on the WHOLE cloud.
Test infrastructure: used by tests/test_gpu_fullsize.py and __graft_entry__.smoke() only.
"""
import numpy as np


def host_xyz_of_chunk(dl, lo, hi):
    """int32 X,Y,Z columns of records [lo, hi) read back from the device record buffer."""
    R = dl.rec_len
    raw = dl.rec[lo * R: hi * R].cpu().numpy()
    dt = np.dtype({"names": ["X", "Y", "Z"], "formats": ["<i4"] * 3, "offsets": [0, 4, 8], "itemsize": R})
    pts = raw.view(dt)
    return {k: np.ascontiguousarray(pts[k]) for k in "XYZ"}


def check_voxel_chunks(dl, res, voxel_size, chunk_size, n_samples=3, seed=0):
    """`res` = device.voxel_downsample(dl, ..., want=("mean","lattice","f32")).  Returns the chunk ids checked."""
    from oracle import las_io, voxel as ov
    n_chunks = -(-dl.n // chunk_size)
    rng = np.random.default_rng(seed)
    ids = sorted(set(rng.choice(n_chunks, size=min(n_samples, n_chunks), replace=False).tolist()) | {n_chunks - 1})
    counts = res.chunk_counts.cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(counts)])
    for c in ids:
        lo, hi = c * chunk_size, min((c + 1) * chunk_size, dl.n)
        las = dict(host_xyz_of_chunk(dl, lo, hi), scales=dl.scales, offsets=dl.offsets, n=hi - lo)
        x, y, z = las_io.scaled(las)
        exp = ov.voxel_down_sample(np.vstack((x, y, z)).T, voxel_size)
        a, b = int(starts[c]), int(starts[c + 1])
        assert b - a == exp.shape[0], f"chunk {c}: {b - a} voxels, oracle {exp.shape[0]}"
        if res.mean is not None:
            assert np.array_equal(res.mean[a:b].cpu().numpy(), exp), f"chunk {c}: voxel means differ from the oracle"
        q = np.stack([las_io.quantise(exp[:, i], dl.scales[i], dl.offsets[i]) for i in range(3)], axis=1)
        if res.lattice is not None:
            assert np.array_equal(res.lattice[a:b].cpu().numpy(), q), f"chunk {c}: re-quantised lattice differs"
        if res.f32 is not None:
            f = np.stack([(q[:, i].astype(np.float64) * dl.scales[i] + dl.offsets[i]) for i in range(3)], axis=1)
            assert np.array_equal(res.f32[a:b].cpu().numpy(), f.astype(np.float32)), f"chunk {c}: float32 cloud differs"
    return ids


def check_ground_whole_cloud(raw_dev, stages):
    """Centroid, percentile base and keep mask of the WHOLE float32 cloud against numpy
    (utils/tower_extraction.py:63-64,81-89).  `stages` = towers.run_stages(raw, want_mask=True)."""
    from oracle import ground
    raw = raw_dev.cpu().numpy()
    cen = np.mean(raw, axis=0)
    assert np.array_equal(cen, stages.centroid), "centroid differs from np.mean"
    z = raw[:, 2] - cen[2]
    mask, base, used = ground.percentile_keep_mask(z)
    assert np.float32(base) == stages.base and used == stages.offset_used, "percentile base differs from numpy"
    got = stages.mask.cpu().numpy().astype(bool)
    assert np.array_equal(got, mask), "keep mask differs from numpy"
    g = int(mask.sum())
    assert stages.filtered.shape[0] == g
    # the filtered cloud itself, on a strided sample of rows (points[mask] in float32)
    idx = np.nonzero(mask)[0]
    take = np.linspace(0, g - 1, num=min(g, 200_000)).astype(np.int64)
    exp = raw[idx[take]] - cen
    assert np.array_equal(stages.filtered[take].cpu().numpy(), exp), "filtered rows differ from numpy"
    return g


def check_dbscan_chunks(filtered_dev, labels_dev, eps=8.0, min_points=80, chunk=50_000, n_samples=5, seed=0):
    """Random DBSCAN chunks against the real scikit-learn (same call as utils/tower_extraction.py:107-112),
    minus the running label offset (= labels handed out by all earlier chunks)."""
    from sklearn.cluster import DBSCAN
    G = filtered_dev.shape[0]
    n_chunks = -(-G // chunk)
    rng = np.random.default_rng(seed)
    ids = sorted(set(rng.choice(n_chunks, size=min(n_samples, n_chunks), replace=False).tolist()) | {n_chunks - 1})
    labels = labels_dev
    checked = []
    for c in ids:
        lo, hi = c * chunk, min((c + 1) * chunk, G)
        pts = filtered_dev[lo:hi].cpu().numpy()
        sk = DBSCAN(eps=eps, min_samples=min_points, n_jobs=-1, algorithm="ball_tree").fit(pts).labels_
        before = labels[:lo]
        off = int(before.max().item()) + 1 if lo > 0 and int(before.max().item()) >= 0 else 0
        exp = np.where(sk >= 0, sk + off, -1).astype(np.int32)
        got = labels[lo:hi].cpu().numpy()
        assert np.array_equal(got, exp), f"DBSCAN chunk {c}: labels differ from scikit-learn"
        checked.append((c, int(sk.max()) + 1))
    return checked
