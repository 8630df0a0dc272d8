"""GPU: the fused pipeline (no LAS file between the stages) equals the reference's two-step flow, and the
host-buffer entry with sliced/overlapped H2D equals the resident-records entry."""
import numpy as np
import pytest

from conftest import make_las_dict

pytestmark = pytest.mark.gpu


def test_pipeline_equals_two_step_oracle(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv, pipeline, synth
    from oracle import las_io, towers as ot, voxel as ov
    n, chunk = 400_000, 100_000
    rec = synth.corridor_records(n, 2, "flat", 22, (0.86, 0.085, 0.005, 0.05))
    las = make_las_dict(rec, synth.SCALES, synth.OFFSETS)
    final, _ = ov.downsample_las_arrays(las, 0.1, chunk)
    q = [las_io.quantise(final[:, i], synth.SCALES[i], synth.OFFSETS[i]) for i in range(3)]
    las2 = dict(las, X=q[0], Y=q[1], Z=q[2], n=len(q[0]))
    inter = {}
    ref = ot.extract_towers_arrays(las2, box="aabb", intermediates=inter)
    dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
    out = pipeline.run_pipeline(dl, 0.1, chunk, box="aabb", keep_stages=True, want_points=True)
    assert out.n_voxels == las2["n"] and out.n_candidates == len(inter["filtered"])
    assert np.array_equal(out.stages.labels.cpu().numpy(), inter["labels"])
    assert [t["label"] for t in out.towers] == [t["label"] for t in ref] and len(ref) >= 1
    for a, b in zip(out.towers, ref):
        assert np.allclose(a["center"], b["center"], atol=1e-4) and np.array_equal(a["points"], b["points"])
    pinned = torch.from_numpy(rec.view(np.uint8).copy()).pin_memory()
    for slice_chunks in (1, 3, 100):
        host = pipeline.run_pipeline_from_host(pinned, n, 34, synth.SCALES, synth.OFFSETS, 0.1, chunk,
                                               slice_chunks=slice_chunks, box="aabb")
        assert (host.n_voxels, host.n_candidates, host.n_clusters) == (out.n_voxels, out.n_candidates, out.n_clusters)
        assert [t["label"] for t in host.towers] == [t["label"] for t in out.towers]
        for a, b in zip(host.towers, out.towers):
            assert np.array_equal(a["center"], b["center"]) and np.array_equal(a["extent"], b["extent"])


def test_host_pack_xyz_entry_equals_full_record_entry(cuda_device):
    """pack="xyz": pageable host records -> 12-byte X,Y,Z stream (host threads) -> H2D -> the same kernels with
    rec_len = 12; results identical to shipping whole 34-byte records."""
    import torch
    from pointcloudhookup_b200 import device as dv, pipeline, synth
    n, chunk = 300_000, 100_000
    rec = synth.corridor_records(n, 2, "hilly", 29, (0.86, 0.085, 0.005, 0.05))
    dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
    ref = pipeline.run_pipeline(dl, 0.1, chunk, box="aabb", keep_stages=True)
    pageable = rec.view(np.uint8).copy()
    pinned = torch.from_numpy(pageable.copy()).pin_memory()
    for src, slice_chunks, threads, raw_every in ((pageable, 1, 1, 0), (pageable, 2, 3, 2), (pageable, 100, 0, 0),
                                                  (pinned, 1, 2, 2), (pinned, 1, 0, 1), (pinned, 1, 0, 3)):
        got = pipeline.run_pipeline_from_host(src, n, 34, synth.SCALES, synth.OFFSETS, 0.1, chunk,
                                              slice_chunks=slice_chunks, pack="xyz", threads=threads,
                                              raw_every=raw_every, box="aabb")
        assert (got.n_voxels, got.n_candidates, got.n_clusters) == (ref.n_voxels, ref.n_candidates, ref.n_clusters)
        assert [t["label"] for t in got.towers] == [t["label"] for t in ref.towers]
        for a, b in zip(got.towers, ref.towers):
            assert np.array_equal(a["center"], b["center"]) and np.array_equal(a["extent"], b["extent"])
    # the 12-byte stream is itself a valid record stream for every decode entry
    lib = dv._native.lib()
    packed = np.zeros(n * 12 + 16, dtype=np.uint8)
    off = (-packed.ctypes.data) % 16
    assert lib.pch_host_pack_xyz(pageable.ctypes.data, n, 34, packed.ctypes.data + off, 2) == 0
    dl12 = dv.upload_records(packed[off: off + n * 12], n, 12, synth.SCALES, synth.OFFSETS)
    assert torch.equal(dv.decode_xyz(dl12, torch.float64), dv.decode_xyz(dl, torch.float64))
    v12 = dv.voxel_downsample(dl12, 0.1, chunk, want=("mean", "z32", "f32"))
    v34 = dv.voxel_downsample(dl, 0.1, chunk, want=("mean", "f32"))
    assert torch.equal(v12.mean, v34.mean) and torch.equal(v12.f32, v34.f32)
    assert torch.equal(v12.z32, v34.f32[:, 2].contiguous())


def test_tile_stream_equals_per_tile_calls(cuda_device):
    """run_tiles_from_host (one tile of look-ahead, two staging buffers) yields, tile by tile, exactly what
    independent calls yield — for tiles of different sizes and record lengths, gathered or whole-record."""
    import torch
    from pointcloudhookup_b200 import device as dv, pipeline, synth
    tiles = []
    for k, (n, seed) in enumerate(((200_000, 51), (90_001, 52), (310_000, 53), (4, 54), (150_000, 55))):
        rec = synth.corridor_records(n, 2, "hilly", seed, (0.86, 0.085, 0.005, 0.05))
        raw = rec.view(np.uint8).reshape(n, 34)
        if k == 2:      # a 36-byte format with two trailing bytes
            raw = np.concatenate([raw, np.full((n, 2), 7, np.uint8)], axis=1)
        tiles.append((np.ascontiguousarray(raw).reshape(-1), n, raw.shape[1]))
    want = []
    for rec, n, rl in tiles:
        dl = dv.upload_records(rec, n, rl, synth.SCALES, synth.OFFSETS)
        want.append(pipeline.run_pipeline(dl, 0.1, 50_000, box="aabb"))
    for pack, pin in (("xyz", False), ("none", True), ("xyz", True)):
        src = [((torch.from_numpy(r).pin_memory() if pin else r), n, rl) for r, n, rl in tiles]
        got = list(pipeline.run_tiles_from_host(iter(src), synth.SCALES, synth.OFFSETS, 0.1, 50_000, slice_chunks=2,
                                                pack=pack, raw_every=3 if pin else 0, box="aabb"))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert (a.n_points, a.n_voxels, a.n_candidates, a.n_clusters) == (b.n_points, b.n_voxels, b.n_candidates, b.n_clusters)
            assert [t["label"] for t in a.towers] == [t["label"] for t in b.towers]
            for x, y in zip(a.towers, b.towers):
                assert np.array_equal(x["center"], y["center"]) and np.array_equal(x["extent"], y["extent"])
    assert list(pipeline.run_tiles_from_host([], synth.SCALES, synth.OFFSETS)) == []


def test_percentile_on_raw_column_and_on_the_fly_compaction_equal_the_literal_order(cuda_device):
    """The select on the RAW z column (side stream) + keep flag derived inside the compaction must equal the
    reference's literal order: shift, percentile of the shifted column, mask, gather."""
    import torch
    from pointcloudhookup_b200 import device as dv, towers as tw
    rng = np.random.default_rng(31)
    for m in (1, 7, 1000, 300_001):
        raw = np.stack([rng.uniform(437000, 437100, m), rng.uniform(3139000, 3139400, m),
                        np.round(rng.gamma(2.0, 4.0, m) + 80, 3)], 1).astype(np.float32)
        d = torch.from_numpy(raw).to(cuda_device)
        filt, cen, base, used, mask = tw.ground_filter_percentile(d, want_mask=True)
        pts = raw - np.mean(raw, axis=0)
        z = pts[:, 2]
        b = np.percentile(z, 25)
        keep = z > b + 3.0
        if keep.sum() < 1000:
            keep = z > b + 1.0
        assert np.array_equal(cen.cpu().numpy(), np.mean(raw, axis=0)) and base == b
        assert np.array_equal(mask.cpu().numpy().astype(bool), keep)
        assert np.array_equal(filt.cpu().numpy(), pts[keep])
        assert torch.equal(dv.f32_column(d, 2), d[:, 2].contiguous())


def test_pipeline_grid_ground_mode_runs_and_matches_self_oracle(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv, towers as tw, synth
    from oracle import ground as og
    n = 300_000
    rec = synth.corridor_records(n, 2, "hilly", 23, (0.86, 0.085, 0.005, 0.05))
    dl = dv.upload_records(rec.view(np.uint8), n, 34, synth.SCALES, synth.OFFSETS)
    raw = dv.voxel_downsample(dl, 0.25, 1_000_000, want=("f32",)).f32
    st = tw.run_stages(raw, ground="grid", want_mask=True, cell=2.0, hag=3.0)
    rawh = raw.cpu().numpy()
    shifted = rawh - np.mean(rawh, axis=0)
    keep, _ = og.grid_min_keep_mask(shifted, 2.0, 3.0)
    assert np.array_equal(st.mask.cpu().numpy().astype(bool), keep)
    assert np.array_equal(st.filtered.cpu().numpy(), shifted[keep])
