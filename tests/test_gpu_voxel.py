"""GPU parity: LAS decode, segmented radix sort and voxel downsample vs the CPU oracle.
Bit-exact for indices/counts/int32 lattice and (because the sums run in input order) for the
float64 means too."""
import numpy as np
import pytest

from conftest import make_las_dict

pytestmark = pytest.mark.gpu


def _corridor(n, towers=2, seed=1, terrain="flat"):
    from pointcloudhookup_b200 import synth
    rec = synth.corridor_records(n, towers, terrain, seed)
    return rec, synth.SCALES, synth.OFFSETS


def test_decode_f64_f32_match_oracle(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv
    from oracle import las_io
    for n in (0, 1, 255, 257, 5000, 70001):
        rec, sc, of = _corridor(n) if n else (np.zeros(0, dtype=_corridor(1)[0].dtype), *_corridor(1)[1:])
        dl = dv.upload_records(rec.view(np.uint8), n, 34, sc, of)
        x64 = dv.decode_xyz(dl, torch.float64).cpu().numpy()
        x32 = dv.decode_xyz(dl, torch.float32).cpu().numpy()
        las = make_las_dict(rec, sc, of)
        ref = np.stack(las_io.scaled(las), axis=1) if n else np.zeros((0, 3))
        assert np.array_equal(x64, ref)
        assert np.array_equal(x32, ref.astype(np.float32))


@pytest.mark.parametrize("rec_len", [20, 26, 28, 34, 37, 63])
def test_decode_other_record_lengths(cuda_device, rec_len):
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(rec_len)
    n = 10007
    raw = rng.integers(0, 256, size=(n, rec_len), dtype=np.uint8)
    xyz = rng.integers(-2**31, 2**31 - 1, size=(n, 3), dtype=np.int64).astype(np.int32)
    raw[:, :12] = xyz.view(np.uint8).reshape(n, 12)
    sc, of = np.array([0.01, 0.001, 0.5]), np.array([1000.0, -5.5, 3.25])
    dl = dv.upload_records(raw.reshape(-1), n, rec_len, sc, of)
    got = dv.decode_xyz(dl, torch.float64).cpu().numpy()
    assert np.array_equal(got, xyz.astype(np.float64) * sc + of)
    mm = dv.chunk_minmax(dl, 3000).cpu().numpy()
    for c in range(mm.shape[0]):
        sl = xyz[c * 3000:(c + 1) * 3000]
        assert np.array_equal(mm[c, :3], sl.min(0)) and np.array_equal(mm[c, 3:], sl.max(0))


@pytest.mark.parametrize("n,seg,lo,hi", [(1, 1, 0, 8), (4097, 4097, 3, 20), (100000, 30000, 17, 50),
                                           (250000, 250000, 0, 64), (123457, 5000, 20, 33)])
def test_segmented_radix_sort(cuda_device, n, seg, lo, hi):
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**63 - 1, size=n, dtype=np.int64)
    if n > 10:
        keys[: n // 3] = keys[0]  # heavy duplicates
    t = torch.from_numpy(keys).to(cuda_device)
    out = dv.sort_u64_segmented(t, seg, lo, hi).cpu().numpy()
    mask = ((1 << (hi - lo)) - 1) if hi - lo < 64 else -1
    exp = np.empty_like(keys)
    for s in range(0, n, seg):
        k = keys[s:s + seg]
        digit = (k.view(np.uint64) >> np.uint64(lo)) & np.uint64(mask & 0xFFFFFFFFFFFFFFFF)
        exp[s:s + seg] = k[np.argsort(digit, kind="stable")]
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("n,chunk,voxel", [(1, 10, 0.1), (1000, 1000, 0.1), (50000, 20000, 0.1),
                                             (300000, 100000, 0.1), (300000, 1000000, 0.25),
                                             (200001, 65536, 1.0)])
def test_voxel_downsample_bit_exact(cuda_device, n, chunk, voxel):
    from pointcloudhookup_b200 import device as dv
    from oracle import las_io, voxel as ov
    rec, sc, of = _corridor(n, towers=2, seed=3)
    dl = dv.upload_records(rec.view(np.uint8), n, 34, sc, of)
    res = dv.voxel_downsample(dl, voxel, chunk, want=("mean", "lattice", "f32"))
    las = make_las_dict(rec, sc, of)
    ref, counts = ov.downsample_las_arrays(las, voxel, chunk)
    assert res.count == ref.shape[0]
    assert np.array_equal(res.chunk_counts.cpu().numpy(), counts)
    got = res.mean.cpu().numpy()
    assert np.array_equal(got, ref), f"max |diff| = {np.abs(got - ref).max()}"
    lat = np.stack([las_io.quantise(ref[:, i], sc[i], of[i]) for i in range(3)], axis=1)
    assert np.array_equal(res.lattice.cpu().numpy(), lat)
    f32 = (lat.astype(np.float64) * sc + of).astype(np.float32)
    assert np.array_equal(res.f32.cpu().numpy(), f32)


def test_encode_roundtrip(cuda_device):
    import torch
    from pointcloudhookup_b200 import device as dv
    rng = np.random.default_rng(5)
    for m in (0, 1, 1023, 1024, 1025, 50000):
        lat = rng.integers(-10**9, 10**9, size=(m, 3)).astype(np.int32)
        rec, mm = dv.encode_records(torch.from_numpy(lat).to(cuda_device), 34)
        raw = rec.cpu().numpy().reshape(m, 34)
        assert np.array_equal(raw[:, :12].copy().view(np.int32).reshape(m, 3), lat)
        assert not raw[:, 12:].any()
        if m:
            assert np.array_equal(mm.cpu().numpy(), np.concatenate([lat.min(0), lat.max(0)]))


def test_reciprocal_division_equals_ieee_divide(cuda_device):
    """The voxel kernels replace `sum/count`, `(mean-offset)/scale` and `(p-origin)/voxel` by a reciprocal
    product + two FMA corrections; on the device that must equal __ddiv_rn bit for bit (pch_selftest_fastdiv
    counts the mismatches) for lattice-like, uniform, decode-then-subtract and random-mantissa operands,
    including the values the range guard turns away (0, denormals, inf, nan)."""
    import ctypes
    import torch
    from pointcloudhookup_b200 import _native
    lib = _native.lib()
    rng = np.random.default_rng(2024)
    n = 4_000_000
    k = rng.integers(-2**31, 2**31, n)
    cases = [k * 0.001 * 0.5,                                       # half-lattice values (the common .5 ties)
             rng.uniform(0, 1e7, n),
             (rng.integers(0, 2**31, n) * 0.001 + 437000.0) - 437000.0,
             (rng.integers(0, 2**31, n) * 0.001 + 3139000.0) * rng.integers(1, 65, n),
             np.ldexp(rng.uniform(1, 2, n), rng.integers(-40, 40, n)) * rng.choice([-1.0, 1.0], n),
             np.array([0.0, -0.0, 5e-324, 1e-310, -1e-250, 1e250, np.inf, -np.inf, np.nan, 1e-199, 1e199] * 1000)]
    bad = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    divisors = [0.001, 0.01, 0.1, 1e-4, 2.5e-4, 0.25, 0.5, 0.3, 1 / 3, 1e-5, 0.05, 1.7, -0.001, 1e-7] + \
               [float(c) for c in range(1, 65)]
    for b in divisors:
        for a in cases:
            d = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(cuda_device)
            rc = lib.pch_selftest_fastdiv(d.data_ptr(), d.numel(), ctypes.c_double(b), bad.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
            assert rc == 0 and int(bad.item()) == 0, (b, int(bad.item()))
    # divisors the kernels route to the true divide are refused here
    allones = np.frombuffer(np.uint64(0x3FEFFFFFFFFFFFFF).tobytes(), dtype=np.float64)[0]
    for b in (0.0, float("nan"), 1e80, 1e-80, float(allones)):
        assert lib.pch_selftest_fastdiv(d.data_ptr(), 10, ctypes.c_double(b), bad.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream) != 0


def test_voxel_downsample_with_ineligible_scale_takes_the_true_divide(cuda_device):
    """A LAS scale whose significand is all ones is not eligible for the reciprocal path: still bit-exact."""
    import torch
    from pointcloudhookup_b200 import device as dv, synth
    from oracle import voxel as ov
    allones = float(np.frombuffer(np.uint64(0x3F4FFFFFFFFFFFFF).tobytes(), dtype=np.float64)[0])   # ~0.000977
    scales = (allones, 0.001, allones)
    n = 120_000
    rec = synth.corridor_records(n, 2, "flat", 5)
    dl = dv.upload_records(rec.view(np.uint8), n, 34, scales, synth.OFFSETS)
    las = {"scales": np.array(scales), "offsets": np.array(synth.OFFSETS), "X": rec["X"].copy(), "Y": rec["Y"].copy(),
           "Z": rec["Z"].copy(), "n": n}
    ref, _ = ov.downsample_las_arrays(las, 0.1, 50_000)
    got = dv.voxel_downsample(dl, 0.1, 50_000, want=("mean", "lattice"))
    assert np.array_equal(got.mean.cpu().numpy(), ref)
    from oracle import las_io
    q = np.stack([las_io.quantise(ref[:, i], scales[i], synth.OFFSETS[i]) for i in range(3)], 1)
    assert np.array_equal(got.lattice.cpu().numpy(), q)
