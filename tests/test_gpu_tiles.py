"""GPU: whole-corridor DBSCAN over spatial tiles with the halo exchange (tiles.py, pipeline.run_pipeline_tiled).
Several ranks are emulated as threads on one device (tiles.ThreadComm: host threads meet at barriers, kernels of
different ranks never wait on each other).  Parity: N tiles + halo == one GPU on the concatenated cloud == the
real scikit-learn on the concatenated cloud, label for label (test/zzzzz.py:79-84)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS, MINPTS = 8.0, 20


def _run_ranks(world, fn):
    from pointcloudhookup_b200 import tiles as tl
    group = tl.ThreadComm.Group(world)
    out, errs = [None] * world, []

    def run(r):
        import torch
        try:
            torch.cuda.set_device(0)
            out[r] = fn(r, tl.ThreadComm(group, r))
        except BaseException as e:
            errs.append(e)
            group.barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_device_tile_dbscan_equals_single_gpu_and_sklearn(cuda_device, world):
    import torch
    from sklearn.cluster import DBSCAN
    import halo_oracle as ho
    from pointcloudhookup_b200 import device as dv, tiles as tl
    tiles = ho.corridor_candidates(5, world, per_tile=12000)
    allp = np.concatenate(tiles)
    ref = DBSCAN(eps=EPS, min_samples=MINPTS, algorithm="ball_tree").fit(allp).labels_.astype(np.int32)
    one = dv.dbscan_chunked(torch.from_numpy(allp).cuda(), EPS, MINPTS, chunk=len(allp))
    assert np.array_equal(one.labels.cpu().numpy(), ref)
    out = _run_ranks(world, lambda r, comm: tl.tile_dbscan(torch.from_numpy(tiles[r]).cuda(), (1.0, 0.0), EPS, MINPTS, comm))
    got = np.concatenate([o.labels.cpu().numpy() for o in out])
    assert np.array_equal(got, ref)
    k = int(ref.max()) + 1
    assert all(o.n_clusters == k == one.n_clusters for o in out)
    assert np.array_equal(out[0].stats["count"], one.stats["count"])
    assert np.array_equal(out[0].stats["min"], one.stats["min"]) and np.array_equal(out[0].stats["max"], one.stats["max"])
    assert np.allclose(out[0].stats["sum"], one.stats["sum"], rtol=1e-9, atol=1e-6)
    for o in out[1:]:
        assert o.stats.tobytes() == out[0].stats.tobytes()
    if world > 1:
        assert sum(sum(o.sent) for o in out) > 0


def test_run_pipeline_tiled_equals_one_gpu_on_the_concatenation(cuda_device):
    """3 tiles of a synthetic corridor on 3 emulated ranks: voxel + the reference's height filter per tile, ONE
    DBSCAN across the tiles.  Same labels, clusters and towers as the single-device run on the concatenated
    candidates; a tower sits right on every cut."""
    import torch
    from pointcloudhookup_b200 import device as dv, pipeline, synth, towers as tw
    world = 3
    axis = pipeline.corridor_axis(synth.AZIMUTH_DEG)
    # one 3-span corridor (sparse conductors, so towers stay separate clusters) cut at s = 525 m — right through
    # the second tower — and at s = 700 m in the open
    rec = synth.corridor_records(1_200_000, 3, "flat", 21, (0.86, 0.085, 0.005, 0.05))
    e = rec["X"] * 0.001 + synth.OFFSETS[0] - synth.ORIGIN_EN[0]
    nn = rec["Y"] * 0.001 + synth.OFFSETS[1] - synth.ORIGIN_EN[1]
    s_along = e * axis[0] + nn * axis[1]
    tile_of = np.digitize(s_along, [525.0, 700.0])
    dls = []
    for r in range(world):
        part = np.ascontiguousarray(rec[tile_of == r])
        assert part.size > 100_000
        dls.append(dv.upload_records(part.view(np.uint8), part.size, 34, synth.SCALES, synth.OFFSETS))
    out = _run_ranks(world, lambda r, comm: pipeline.run_pipeline_tiled([dls[r]], comm, axis, 0.1, 100_000, EPS, 80, keep=True))
    origin = out[0].origin
    assert all(np.array_equal(o.origin, origin) for o in out)
    allc = torch.cat([o.candidates for o in out]).contiguous()
    one = dv.dbscan_chunked(allc, EPS, 80, chunk=allc.shape[0])
    got = torch.cat([o.labels for o in out])
    assert torch.equal(got, one.labels)
    assert all(o.n_clusters == one.n_clusters for o in out)
    stages = tw.TowerStages(None, origin, np.float32("nan"), 3.0, allc, one.labels, one.n_clusters, one.stats)
    ref_towers = tw.select_towers(stages, box="aabb", want_points=False)
    for o in out:
        assert [t["label"] for t in o.towers] == [t["label"] for t in ref_towers]
        for a, b in zip(o.towers, ref_towers):
            assert np.allclose(a["center"], b["center"], atol=1e-4) and np.allclose(a["extent"], b["extent"], atol=1e-4)
    assert len(ref_towers) >= 2
    # the same result with ONE rank holding all three tiles (a GPU with several tiles)
    solo = pipeline.run_pipeline_tiled(dls, __import__("pointcloudhookup_b200.tiles", fromlist=["SoloComm"]).SoloComm(), axis,
                                       0.1, 100_000, EPS, 80, keep=True)
    assert torch.equal(solo.labels, one.labels) and [t["label"] for t in solo.towers] == [t["label"] for t in ref_towers]
    # some cluster really spans a cut
    offs = np.concatenate([[0], np.cumsum([o.n_candidates for o in out])])
    lab = one.labels.cpu().numpy()
    spans = 0
    for c in range(one.n_clusters):
        idx = np.nonzero(lab == c)[0]
        spans = max(spans, len({int(np.searchsorted(offs, i, side="right")) for i in idx[:: max(1, len(idx) // 512)]}))
    assert spans >= 2
