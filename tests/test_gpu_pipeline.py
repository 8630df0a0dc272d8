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
