"""GPU parity of the tower-extraction stages against the oracle (real numpy percentile, real
scikit-learn DBSCAN): centroid, height filter, all_labels and the tower list, bit-exact for
labels/masks/counts, 1e-4 m for float box outputs."""
import numpy as np
import pytest

from conftest import make_las_dict

pytestmark = pytest.mark.gpu


def _downsampled_f32(n, towers, seed, terrain="flat", fractions=(0.80, 0.08, 0.07, 0.05), voxel=0.1, chunk=500000):
    """Oracle-side: synthetic corridor -> voxel downsample -> LAS re-quantisation -> dict."""
    from pointcloudhookup_b200 import synth
    from oracle import las_io, voxel as ov
    rec = synth.corridor_records(n, towers, terrain, seed, fractions)
    las = make_las_dict(rec, synth.SCALES, synth.OFFSETS)
    final, _ = ov.downsample_las_arrays(las, voxel, chunk)
    sc, of = synth.SCALES, synth.OFFSETS
    q = [las_io.quantise(final[:, i], sc[i], of[i]) for i in range(3)]
    las2 = dict(las, X=q[0], Y=q[1], Z=q[2], n=len(q[0]))
    return rec, las2


def test_centroid_sequential_f32_bit_exact(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(7)
    for m in (1, 7, 8, 9, 1000, 2047, 2048, 2049, 250001, 3_000_000):
        a = np.stack([437000 + rng.random(m) * 3000, 3.139e6 + rng.random(m) * 3000, 80 + rng.random(m) * 40],
                     axis=1).astype(np.float32)
        t = torch.from_numpy(a).to(cuda_device)
        for serial in (True, False):
            cen, sums = dv.f32_centroid(t, serial=serial)
            assert np.array_equal(sums.cpu().numpy(), np.add.reduce(a, axis=0)), (m, serial)
            assert np.array_equal(cen.cpu().numpy(), np.mean(a, axis=0)), (m, serial)


def test_centroid_parallel_exact_on_adversarial_columns(cuda_device):
    """The binade-map evaluation must equal numpy's sequential float32 sum for ties, huge dynamic range,
    stagnation (sum stops growing), zeros, denormals, negative and mixed-sign data (real-add fallback)."""
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(11)
    m = 700_000
    cols = {
        "ties": rng.integers(0, 64, m) * 0.5,
        "wide": 10 ** rng.uniform(-3, 6, m),
        "tiny": rng.random(m) * 1e-3,
        "const": np.full(m, 437500.0),
        "zeros_mixed": np.where(rng.random(m) < 0.5, 0.0, rng.random(m) * 100),
        "denormal": rng.random(m) * 1e-40,
        "negative": -(3.139e6 + rng.random(m) * 3000),
        "mixed_sign": rng.normal(0, 50, m),
        "spike": np.where(np.arange(m) == 400_000, 3.0e12, rng.random(m) * 10),
    }
    names = list(cols)
    for i in range(0, len(names), 3):
        trio = names[i:i + 3]
        a = np.stack([cols[k] for k in trio], axis=1).astype(np.float32)
        cen, sums, st = dv.f32_centroid(torch.from_numpy(a).to(cuda_device), want_stats=True)
        exp = np.add.reduce(a, axis=0)
        assert np.array_equal(sums.cpu().numpy(), exp), (trio, sums.cpu().numpy(), exp, st)
        assert np.array_equal(cen.cpu().numpy(), np.mean(a, axis=0)), trio
    # stagnation regime: 30 M points, the float32 sum stops growing long before the end
    big = 20_000_000
    a = np.stack([437000 + rng.random(big) * 3000, 3.139e6 + rng.random(big) * 3000, 80 + rng.random(big) * 40],
                 axis=1).astype(np.float32)
    cen, sums, st = dv.f32_centroid(torch.from_numpy(a).to(cuda_device), want_stats=True)
    assert np.array_equal(sums.cpu().numpy(), np.add.reduce(a, axis=0)), st
    assert st[:, 1].max() < 200, st     # almost every tile went through the composed maps


@pytest.mark.parametrize("n", [1, 2, 3, 10, 1001, 65537, 300000])
def test_percentile_select_and_lerp(cuda_device, n):
    import torch
    from pointcloudhookup_b200 import device as dv, towers as tw
    rng = np.random.default_rng(n)
    z = (rng.normal(0, 5, n)).astype(np.float32)
    if n > 10:
        z[: n // 2] = z[0]  # ties
    for q in (25, 10, 20, 50, 0, 100):
        r0, r1, gamma = tw.percentile_ranks_f32(n, q)
        two = dv.select_f32(torch.from_numpy(z).to(cuda_device), r0, r1).cpu().numpy()
        srt = np.sort(z)
        assert two[0] == srt[r0] and two[1] == srt[r1]
        got = tw.percentile_lerp_f32(two[0], two[1], gamma)
        exp = np.percentile(z, q)
        assert got.dtype == exp.dtype == np.float32 and got == exp


def _check_stages(stages, inter):
    assert np.array_equal(stages.centroid, inter["centroid"])
    assert stages.base == inter["base"] and stages.offset_used == inter["offset_used"]
    assert np.array_equal(stages.filtered.cpu().numpy(), inter["filtered"])
    got = stages.labels.cpu().numpy()
    assert np.array_equal(got, inter["labels"]), f"{(got != inter['labels']).sum()} labels differ"
    k = int(inter["labels"].max()) + 1 if inter["labels"].size else 0
    assert stages.n_clusters == k
    for lab in range(k):
        pts = inter["filtered"][inter["labels"] == lab]
        st = stages.stats[lab]
        assert st["count"] == len(pts)
        assert np.array_equal(st["min"], pts.min(0)) and np.array_equal(st["max"], pts.max(0))
        assert np.allclose(st["sum"] / st["count"], pts.astype(np.float64).mean(0), atol=1e-6)


@pytest.mark.parametrize("n,towers,seed,fractions", [
    (300000, 2, 11, (0.80, 0.08, 0.07, 0.05)),          # conductors dense: towers chained by wires
    (600000, 3, 12, (0.86, 0.085, 0.005, 0.05)),        # sparse conductors: isolated tower clusters
    (1000000, 5, 1, (0.80, 0.08, 0.07, 0.05)),          # BASELINE config 0
])
def test_tower_stages_match_oracle(cuda_device, n, towers, seed, fractions):
    import torch
    from pointcloudhookup_b200 import towers as tw
    from oracle import towers as ot
    _, las2 = _downsampled_f32(n, towers, seed, fractions=fractions)
    inter = {}
    ref_aabb = ot.extract_towers_arrays(las2, box="aabb", intermediates=inter)
    raw = torch.from_numpy(inter["raw"]).to(cuda_device)
    stages = tw.run_stages(raw)
    _check_stages(stages, inter)
    got = tw.select_towers(stages, box="aabb")
    assert [t["label"] for t in got] == [t["label"] for t in ref_aabb]
    for a, b in zip(got, ref_aabb):
        assert np.allclose(a["center"], b["center"], atol=1e-4) and np.allclose(a["extent"], b["extent"], atol=1e-4)
        assert np.array_equal(a["points"], b["points"])
    ref_obb = ot.extract_towers_arrays(las2, box="obb")
    got_obb = tw.select_towers(stages, box="obb")
    assert [t["label"] for t in got_obb] == [t["label"] for t in ref_obb]
    for a, b in zip(got_obb, ref_obb):
        assert np.allclose(a["center"], b["center"], atol=1e-4) and np.allclose(a["extent"], b["extent"], atol=1e-4)
        assert abs(a["north_angle"] - b["north_angle"]) < 1e-6


def test_dbscan_exact_eps_ties_and_borders(cuda_device):
    """Lattice points exactly eps apart are neighbours (d <= eps), border points go to the
    lowest-numbered cluster, labels are ordered by first core index; several ragged chunks."""
    import torch
    from sklearn.cluster import DBSCAN
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(3)
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([g * 2.0, g * 2.0 + np.array([60.0, 0, 0]), rng.uniform(-5, 90, (400, 3))]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    for eps, ms, chunk in ((2.0, 7, 1000000), (4.0, 30, 300), (2.0, 5, 257), (8.0, 80, 500)):
        res = dv.dbscan_chunked(torch.from_numpy(pts).to(cuda_device), eps, ms, chunk)
        exp = np.full(len(pts), -1, np.int32)
        cur = 0
        for s in range(0, len(pts), chunk):
            lab = DBSCAN(eps=eps, min_samples=ms, algorithm="ball_tree").fit(pts[s:s + chunk]).labels_
            lab[lab != -1] += cur
            exp[s:s + chunk] = lab
            cur = lab.max() + 1 if (lab != -1).any() else cur
        assert np.array_equal(res.labels.cpu().numpy(), exp), (eps, ms, chunk)
        assert res.n_clusters == cur


def test_grid_min_ground_matches_self_oracle(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv
    from oracle import ground as og
    rng = np.random.default_rng(9)
    p = np.stack([rng.uniform(0, 300, 200000), rng.uniform(0, 60, 200000), rng.normal(50, 3, 200000)], 1).astype(np.float32)
    keep, gz = dv.grid_min_ground(torch.from_numpy(p).to(cuda_device), 2.0, 3.0)
    ek, egz = og.grid_min_keep_mask(p, 2.0, 3.0)
    assert np.array_equal(gz.cpu().numpy(), egz)
    assert np.array_equal(keep.cpu().numpy().astype(bool), ek)


def test_dbscan_randomised_against_sklearn(cuda_device):
    """Random clouds, random eps / min_samples / chunk sizes (ragged last chunk, chunks smaller than a tile, a
    single chunk), mixtures of dense blobs, sheets, lines and uniform noise: labels must equal scikit-learn's
    chunk by chunk, including the running label offset."""
    import torch
    from sklearn.cluster import DBSCAN
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(2024)
    for trial in range(24):
        n = int(rng.integers(200, 60000))
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.integers(50, max(60, n // 3)))
            kind = rng.integers(0, 4)
            c = rng.uniform(-200, 200, 3)
            if kind == 0:
                parts.append(c + rng.normal(0, rng.uniform(0.5, 6), (k, 3)))                       # blob
            elif kind == 1:
                parts.append(c + rng.uniform(-40, 40, (k, 3)) * np.array([1, 1, 0.01]))            # sheet
            elif kind == 2:
                t = rng.uniform(0, 1, (k, 1))
                parts.append(c + t * rng.uniform(-150, 150, 3) + rng.normal(0, 0.2, (k, 3)))        # line
            else:
                parts.append(rng.uniform(-250, 250, (k, 3)))                                        # noise
        pts = np.concatenate(parts)[:n].astype(np.float32)
        pts = pts[rng.permutation(len(pts))]
        if trial % 5 == 0:
            pts = np.round(pts * 2) / 2            # lattice data: exact distance ties
        eps = float(rng.choice([0.75, 2.0, 3.5, 8.0, 10.0]))
        ms = int(rng.choice([3, 5, 20, 50, 80]))
        chunk = int(rng.choice([len(pts), 50000, 4097, 1500, 997]))
        res = dv.dbscan_chunked(torch.from_numpy(pts).to(cuda_device), eps, ms, chunk)
        exp = np.full(len(pts), -1, np.int32)
        cur = 0
        for s in range(0, len(pts), chunk):
            lab = DBSCAN(eps=eps, min_samples=ms, algorithm="ball_tree").fit(pts[s:s + chunk]).labels_
            lab[lab != -1] += cur
            exp[s:s + chunk] = lab
            cur = lab.max() + 1 if (lab != -1).any() else cur
        got = res.labels.cpu().numpy()
        assert np.array_equal(got, exp), (trial, n, eps, ms, chunk, int((got != exp).sum()))
        assert res.n_clusters == cur


def test_ground_filter_variants_of_the_reference_scripts(cuda_device):
    """The height-filter variants of the scratch scripts are the same kernel path with other parameters:
    pct10 + 4 (test/main_ground.py:118-131), pct20 + 2.5 (test/008.py:212-224), min(z) + 6 (test/zzzzz.py:64-65,
    = percentile 0), plus the extreme q = 100; each against numpy on the shifted float32 column."""
    import torch
    from pointcloudhookup_b200 import towers as tw
    rng = np.random.default_rng(77)
    m = 400_003
    raw = np.stack([rng.uniform(437000, 437500, m), rng.uniform(3139000, 3139300, m),
                    np.round(80 + rng.gamma(2.0, 3.0, m), 3)], 1).astype(np.float32)
    d = torch.from_numpy(raw).to(cuda_device)
    pts = raw - np.mean(raw, axis=0)
    z = pts[:, 2]
    for pct, off in ((10, 4.0), (20, 2.5), (0, 6.0), (25, 3.0), (50, 0.0), (100, -1.0), (33.3, 1.25)):
        filt, cen, base, used, mask = tw.ground_filter_percentile(d, pct=pct, offset=off, min_keep=0, want_mask=True)
        b = np.percentile(z, pct)
        keep = z > b + off
        assert base == b and used == off, (pct, float(base), float(b))
        assert np.array_equal(mask.cpu().numpy().astype(bool), keep), pct
        assert np.array_equal(filt.cpu().numpy(), pts[keep]), pct


def test_grid_float32_reciprocal_division_equals_ieee_divide(cuda_device):
    """The grid min-z kernels evaluate (p - min) / cell as a reciprocal product + FMA corrections (float32); on the
    device that must equal __fdiv_rn bit for bit for shifted coordinates, random mantissas and the values the range
    guard turns away."""
    import ctypes
    import torch
    from pointcloudhookup_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(77)
    n = 4_000_000
    cases = [rng.uniform(0, 2e4, n).astype(np.float32),
             (rng.uniform(437000, 438000, n).astype(np.float32) - np.float32(437000.0)),
             (rng.integers(0, 2**24, n) * np.float32(0.5)).astype(np.float32),
             np.ldexp(rng.uniform(1, 2, n), rng.integers(-30, 30, n)).astype(np.float32) * rng.choice([-1.0, 1.0], n).astype(np.float32),
             np.array([0.0, -0.0, 1e-45, 1e-39, -1e-32, 1e32, np.inf, -np.inf, np.nan] * 1000, dtype=np.float32)]
    bad = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    for b in [2.0, 0.5, 1.0, 3.0, 0.1, 0.25, 1.5, 0.3, 2.5, 5.0, 10.0, 0.7, 1 / 3, 7.3]:
        for a in cases:
            d = torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
            rc = lib.pch_selftest_fastdiv_f32(d.data_ptr(), d.numel(), ctypes.c_float(b), bad.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.pch_last_error()
            assert int(bad.item()) == 0, (b, a[:4])
