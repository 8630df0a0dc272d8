"""GPU: SURVEY §8f-3 — all-towers box crop (test/kuangxuan.py:58-79) and the preview subsample
(pyGUI_towers_test.py:174-177) against the numpy restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(n, seed, rec_len=34):
    from pointcloudhookup_b200 import synth
    rec = synth.corridor_records(n, 3, "hilly", seed, (0.80, 0.08, 0.07, 0.05))
    if rec_len == 34:
        return rec.view(np.uint8).reshape(-1), rec
    raw = np.zeros((n, rec_len), dtype=np.uint8)
    raw[:, :12] = rec.view(np.uint8).reshape(n, 34)[:, :12]
    raw[:, 12:] = 0xAB
    return raw.reshape(-1), rec


@pytest.mark.parametrize("rec_len", [34, 20, 12, 37])
def test_crop_boxes_equal_the_reference_masks(cuda_device, rec_len):
    import torch
    from pointcloudhookup_b200 import crop, device as dv, synth
    from oracle import crop as oc, las_io
    n = 300_000
    raw, rec = _cloud(n, 41, rec_len)
    dl = dv.upload_records(raw, n, rec_len, synth.SCALES, synth.OFFSETS)
    pts = dv.decode_xyz(dl, torch.float64).cpu().numpy()
    las = {"scales": synth.SCALES, "offsets": synth.OFFSETS, "X": rec["X"], "Y": rec["Y"], "Z": rec["Z"], "n": n}
    x, y, z = las_io.scaled(las)
    assert np.array_equal(pts, np.vstack((x, y, z)).T)
    rng = np.random.default_rng(5)
    towers = []
    for i in rng.choice(n, 70, replace=False):           # > 64 boxes: two launches
        towers.append({"x": pts[i, 0], "y": pts[i, 1], "z": pts[i, 2] - 10, "width": rng.uniform(5, 25),
                       "height": rng.uniform(10, 40)})
    boxes = np.stack([oc.kuangxuan_bounds(t) for t in towers])
    boxes[3] = boxes[2]                                   # identical boxes
    boxes[5, 3] = boxes[5, 0] - 1                         # empty box (max < min)
    boxes[7] = [pts[11, 0], pts[11, 1], pts[11, 2], pts[11, 0], pts[11, 1], pts[11, 2]]   # degenerate: inclusive bounds
    boxes[9] = [-np.inf, -np.inf, -np.inf, np.inf, np.inf, np.inf]                        # the whole cloud
    got = crop.crop_boxes(dl, boxes)
    want = oc.crop_boxes(pts, boxes)
    assert len(got) == len(want) == 70
    for b, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g.cpu().numpy(), w), b
    assert got[5].shape[0] == 0 and got[7].shape[0] >= 1 and got[9].shape[0] == n
    # dict entry = the reference's loop
    got2 = crop.crop_towers(dl, towers[:4])
    for t, g in zip(towers[:4], got2):
        assert np.array_equal(g.cpu().numpy(), oc.crop_boxes(pts, oc.kuangxuan_bounds(t)[None])[0])
    assert crop.crop_towers(dl, []) == []
    e = dv.upload_records(np.zeros(0, np.uint8), 0, 34, synth.SCALES, synth.OFFSETS)
    assert [t.shape for t in crop.crop_boxes(e, boxes[:2])] == [(0, 3), (0, 3)]


def test_preview_subsample_draws_distinct_points(cuda_device):
    import torch
    from pointcloudhookup_b200 import crop, device as dv, synth
    from oracle import crop as oc
    n = 250_001
    raw, rec = _cloud(n, 43)
    dl = dv.upload_records(raw, n, 34, synth.SCALES, synth.OFFSETS)
    pts = dv.decode_xyz(dl, torch.float64).cpu().numpy()
    for k, seed in ((200_000, 7), (200_000, None), (1000, 123456789012345), (250_001, 9), (1, 3)):
        idx = crop.sample_indices(n, k, seed, cuda_device).cpu().numpy()
        assert idx.shape == (k,) and idx.min() >= 0 and idx.max() < n and np.unique(idx).size == k
        assert np.array_equal(idx.astype(np.uint64), oc.sample_indices(n, k, seed))
        if k < n:
            assert np.array_equal(crop.preview_subsample(dl, k, seed).cpu().numpy(), pts[idx])
    assert np.array_equal(crop.preview_subsample(dl, 300_000, 1).cpu().numpy(), pts)     # small cloud: untouched
    a = crop.sample_indices(n, 50_000, 11, cuda_device).cpu().numpy()
    b = crop.sample_indices(n, 50_000, 12, cuda_device).cpu().numpy()
    assert not np.array_equal(a, b)
    # spread: a keyed sample of 50k of 250k covers every decile of the index range about evenly
    h = np.histogram(a, bins=10, range=(0, n))[0]
    assert h.min() > 4000 and h.max() < 6000
