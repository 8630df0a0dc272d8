"""Device-side orchestration: torch supplies HBM buffers, pinned staging and the CUDA stream;
every per-point operation is a kernel of libpch_b200.so called through the C ABI (_native).

Nothing here computes on the CPU and nothing falls back: without a CUDA device or without the
built library these functions raise.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import threading
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native
from ._native import VoxelPlan, check, d3


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _native.NativeError("no CUDA device: pointcloudhookup_b200 has no CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclasses.dataclass
class DeviceLas:
    """Raw LAS point records resident in HBM (record i at byte i*rec_len, buffer padded to 16 B)."""
    rec: torch.Tensor           # uint8, >= align16(n*rec_len) bytes
    n: int
    rec_len: int
    scales: np.ndarray
    offsets: np.ndarray

    @property
    def device(self):
        return self.rec.device


def padded_bytes(n: int, rec_len: int) -> int:
    return (n * rec_len + 15) // 16 * 16 + 16


def upload_records(records, n: int, rec_len: int, scales, offsets, device=None,
                   pinned: Optional[torch.Tensor] = None) -> DeviceLas:
    """Host record bytes -> HBM.  `records` is a uint8 numpy array / memmap (or a pinned uint8
    tensor); the copy goes through pinned memory so it is a real async DMA on the current stream."""
    _require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    nbytes = n * rec_len
    dev = torch.empty(padded_bytes(n, rec_len), dtype=torch.uint8, device=device)
    if nbytes:
        if isinstance(records, torch.Tensor):
            host = records.view(torch.uint8).reshape(-1)[:nbytes]
            if not host.is_pinned():
                host = host.pin_memory()
        else:
            if pinned is None or pinned.numel() < nbytes:
                pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host = pinned[:nbytes]
            host.numpy()[:] = np.asarray(records).view(np.uint8).reshape(-1)[:nbytes]
        dev[:nbytes].copy_(host, non_blocking=True)
    dev[nbytes:].zero_()
    return DeviceLas(dev, int(n), int(rec_len), np.asarray(scales, dtype=np.float64),
                     np.asarray(offsets, dtype=np.float64))


# Pinned staging is expensive to create (~0.3 s/GB) and must not accumulate per worker thread (the reference GUI
# starts a new thread per action): one locked, size-bounded pool serves every host->device path.  A buffer is owned
# by exactly one transfer between acquire and release.
_POOL = []
_POOL_LOCK = threading.Lock()
_POOL_KEEP = 4


def _alloc_pinned(nbytes: int) -> torch.Tensor:
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def acquire_staging(nbytes: int) -> torch.Tensor:
    nbytes = max(int(nbytes), 16)
    with _POOL_LOCK:
        fit = [i for i, b in enumerate(_POOL) if b.numel() >= nbytes]
        if fit:
            return _POOL.pop(min(fit, key=lambda i: _POOL[i].numel()))     # by index: tensors compare elementwise
    return _alloc_pinned(nbytes)


def release_staging(buf) -> None:
    if buf is None:
        return
    with _POOL_LOCK:
        if any(b is buf for b in _POOL):
            return
        _POOL.append(buf)
        _POOL.sort(key=lambda b: -b.numel())
        del _POOL[_POOL_KEEP:]


def host_threads() -> int:
    """Host threads for the staging gather: this process's share of the cores it may run on."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, cores // max(1, local_world))


def upload_records_xyz(records, n: int, rec_len: int, scales, offsets, device=None,
                       block_points: int = 1 << 22, threads: int = 0) -> DeviceLas:
    """Host record bytes (ndarray, np.memmap of the LAS file, or tensor; pageable is fine) -> HBM as the
    dense 12-byte X,Y,Z stream (rec_len = 12), double-buffered: host threads gather block i+1 into pinned
    staging (pch_host_pack_xyz; page-cache reads of a memmap happen inside those threads) while block i is
    still crossing PCIe.  Everything the drop-in modules compute reads only X,Y,Z, so the other
    rec_len-12 bytes of each record never leave the host."""
    _require_cuda()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    lib = _native.lib()
    dev = torch.empty(padded_bytes(n, 12), dtype=torch.uint8, device=device)
    dev[n * 12:].zero_()
    out = DeviceLas(dev, int(n), 12, np.asarray(scales, dtype=np.float64), np.asarray(offsets, dtype=np.float64))
    if n == 0:
        return out
    if isinstance(records, torch.Tensor):
        keep = records.view(torch.uint8).reshape(-1)
        base, have = keep.data_ptr(), keep.numel()
    else:
        keep = np.asarray(records).view(np.uint8).reshape(-1)
        base, have = keep.ctypes.data, keep.size
    if have < n * rec_len:
        raise ValueError("record buffer shorter than n * rec_len")
    block = max(4, min(int(block_points), n + 3) // 4 * 4)
    stage = [acquire_staging(block * 12) for _ in range(2)]
    nt = threads or host_threads()
    busy = [None, None]
    try:
        with torch.cuda.device(device):
            for bi, lo in enumerate(range(0, n, block)):
                cnt = min(block, n - lo)
                b = bi & 1
                if busy[b] is not None:
                    busy[b].synchronize()          # the copy that last read this staging block has finished
                check(lib.pch_host_pack_xyz(base + lo * rec_len, cnt, rec_len, stage[b].data_ptr(), nt), "pch_host_pack_xyz")
                dev[lo * 12: (lo + cnt) * 12].copy_(stage[b][: cnt * 12], non_blocking=True)
                busy[b] = torch.cuda.Event()
                busy[b].record()
    finally:
        for e in busy:
            if e is not None:
                e.synchronize()                 # the blocks go back to the pool only once their copies are done
        for b in stage:
            release_staging(b)
    return out


# ------------------------------------------------------------------------------------------------
def decode_xyz(dl: DeviceLas, dtype=torch.float64) -> torch.Tensor:
    """(n,3) float64 `np.vstack((las.x,las.y,las.z)).T` or float32 `.astype(np.float32)`."""
    out = torch.empty((dl.n, 3), dtype=dtype, device=dl.device)
    fn = _native.lib().pch_las_decode_f64 if dtype == torch.float64 else _native.lib().pch_las_decode_f32
    check(fn(dl.rec.data_ptr(), dl.n, dl.rec_len, d3(dl.scales), d3(dl.offsets), out.data_ptr(), _stream()),
          "pch_las_decode")
    return out


def quantise(xyz: torch.Tensor, scales, offsets) -> torch.Tensor:
    assert xyz.dtype == torch.float64 and xyz.is_contiguous()
    m = xyz.shape[0]
    out = torch.empty((m, 3), dtype=torch.int32, device=xyz.device)
    check(_native.lib().pch_las_quantise(xyz.data_ptr(), m, d3(scales), d3(offsets), out.data_ptr(), _stream()),
          "pch_las_quantise")
    return out


def encode_records(lattice: torch.Tensor, rec_len: int):
    """(m,3) int32 -> (uint8 records tensor, int32[6] lattice min/max)."""
    assert lattice.dtype == torch.int32 and lattice.is_contiguous()
    m = lattice.shape[0]
    out = torch.empty(padded_bytes(m, rec_len), dtype=torch.uint8, device=lattice.device)
    mm = torch.empty(6, dtype=torch.int32, device=lattice.device)
    check(_native.lib().pch_las_encode(lattice.data_ptr(), m, rec_len, out.data_ptr(), mm.data_ptr(), _stream()),
          "pch_las_encode")
    return out[: m * rec_len], mm


def chunk_minmax(dl: DeviceLas, chunk_size: int, xyz16: Optional[torch.Tensor] = None) -> torch.Tensor:
    cs = max(1, min(int(chunk_size), max(dl.n, 1)))
    n_chunks = max(1, -(-dl.n // cs))
    mm = torch.empty((n_chunks, 6), dtype=torch.int32, device=dl.device)
    check(_native.lib().pch_las_chunk_minmax(dl.rec.data_ptr(), dl.n, dl.rec_len, cs, mm.data_ptr(), _ptr(xyz16),
                                             _stream()), "pch_las_chunk_minmax")
    return mm


def sort_u64_segmented(keys: torch.Tensor, seg_size: int, bit_lo: int, bit_hi: int,
                       tmp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns the tensor (keys or tmp) that holds the sorted result."""
    assert keys.dtype == torch.int64 and keys.is_contiguous()
    n = keys.numel()
    if n == 0 or bit_hi <= bit_lo:
        return keys
    lib = _native.lib()
    if tmp is None:
        tmp = torch.empty_like(keys)
    ws_bytes = lib.pch_sort_workspace_bytes(n, seg_size, bit_lo, bit_hi)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    check(lib.pch_sort_u64_segmented(keys.data_ptr(), tmp.data_ptr(), n, seg_size, bit_lo, bit_hi,
                                     ws.data_ptr(), ws_bytes, _stream()), "pch_sort_u64_segmented")
    n_passes = (bit_hi - bit_lo + 7) // 8
    return tmp if n_passes % 2 else keys


@dataclasses.dataclass
class VoxelResult:
    count: int                              # M
    chunk_counts: torch.Tensor              # int64 [n_chunks]
    mean: Optional[torch.Tensor] = None     # (M,3) float64, canonical (ix,iy,iz) order per chunk
    lattice: Optional[torch.Tensor] = None  # (M,3) int32 re-quantised
    f32: Optional[torch.Tensor] = None      # (M,3) float32 of the re-quantised values
    plan: Optional[dict] = None
    sorted_keys: Optional[torch.Tensor] = None
    z32: Optional[torch.Tensor] = None      # (M,) the z column of f32 as a dense array
    chunk_minmax: Optional[torch.Tensor] = None   # (n_chunks,6) int32 lattice extrema of the INPUT points per chunk


class VoxelSink:
    """Preallocated float32 outputs that several voxel_downsample calls append to (the sliced host-buffer
    pipeline): rows [0, count) are filled."""

    def __init__(self, capacity: int, device, want_z: bool = True):
        self.f32 = torch.empty((int(capacity), 3), dtype=torch.float32, device=device)
        self.z32 = torch.empty(int(capacity), dtype=torch.float32, device=device) if want_z else None
        self.count = 0


def voxel_downsample(dl: DeviceLas, voxel_size: float, chunk_size: int,
                     want: Sequence[str] = ("mean",), keep_keys: bool = False,
                     sink: Optional[VoxelSink] = None) -> VoxelResult:
    """open3d voxel_down_sample over consecutive chunk_size-point slices of the records
    (ui/import_PC.py:45-60), entirely on the device.  `want` selects the outputs ("mean", "lattice",
    "f32", "z32"); with a `sink` the float32 outputs are appended to its buffers instead."""
    _require_cuda()
    if voxel_size <= 0:
        raise ValueError("voxel_size must be > 0")
    lib = _native.lib()
    dev = dl.device
    n = dl.n
    cs = max(1, min(int(chunk_size), max(n, 1)))
    n_chunks = max(1, -(-n // cs))
    if n == 0:
        z = lambda dt: torch.zeros((0, 3), dtype=dt, device=dev)
        return VoxelResult(0, torch.zeros(n_chunks, dtype=torch.int64, device=dev),
                           z(torch.float64) if "mean" in want else None,
                           z(torch.int32) if "lattice" in want else None,
                           z(torch.float32) if "f32" in want else None)
    st = _stream()
    sc, of = d3(dl.scales), d3(dl.offsets)
    # worst case every point is its own voxel; outputs are sized n and sliced after the count is known
    mean = torch.empty((n, 3), dtype=torch.float64, device=dev) if "mean" in want else None
    lat = torch.empty((n, 3), dtype=torch.int32, device=dev) if "lattice" in want else None
    if sink is not None:
        if sink.f32.shape[0] - sink.count < n:
            raise ValueError("VoxelSink too small for this slice")
        f32, z32 = sink.f32[sink.count:], (sink.z32[sink.count:] if sink.z32 is not None else None)
    else:
        f32 = torch.empty((n, 3), dtype=torch.float32, device=dev) if "f32" in want else None
        z32 = torch.empty(n, dtype=torch.float32, device=dev) if "z32" in want else None
    xyz16 = torch.empty((n, 4), dtype=torch.int32, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    tmp = torch.empty(n, dtype=torch.int64, device=dev)
    mm = torch.empty((n_chunks, 6), dtype=torch.int32, device=dev)
    origins = torch.empty((n_chunks, 3), dtype=torch.float64, device=dev)
    plan_dev = torch.empty(8, dtype=torch.int32, device=dev)
    counts = torch.empty(n_chunks, dtype=torch.int64, device=dev)
    ws_bytes = lib.pch_voxel_downsample_las_workspace_bytes(n, cs)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    # one call, no host round trip inside: extrema -> plan (device) -> keys + digit histograms -> radix passes -> reduce
    check(lib.pch_voxel_downsample_las(dl.rec.data_ptr(), n, dl.rec_len, cs, sc, of, float(voxel_size),
                                       xyz16.data_ptr(), keys.data_ptr(), tmp.data_ptr(), mm.data_ptr(),
                                       origins.data_ptr(), plan_dev.data_ptr(), _ptr(mean), _ptr(lat), _ptr(f32), _ptr(z32),
                                       counts.data_ptr(), ws.data_ptr(), ws_bytes, st), "pch_voxel_downsample_las")
    head = ws[:64].cpu().numpy()            # the stage's only device->host read: error word, M, plan
    m = int(head[8:16].view(np.int64)[0])
    plan = VoxelPlan(*[int(v) for v in head[32:64].view(np.int32)])
    if int(head[:4].view(np.int32)[0]):
        raise _native.NativeError("device look-back spin limit hit in voxel_downsample")
    if m < 0 or plan.status != 0:
        del xyz16, keys, tmp, mean, lat
        return _voxel_downsample_wide(dl, voxel_size, cs, want, sink)
    if sink is not None:
        sink.count += m
    skeys = None
    if keep_keys:
        skeys = tmp if plan.n_passes % 2 else keys
    return VoxelResult(m, counts,
                       mean[:m] if mean is not None else None,
                       lat[:m] if lat is not None else None,
                       f32[:m] if f32 is not None else None,
                       plan={k: getattr(plan, k) for k, _ in VoxelPlan._fields_},
                       sorted_keys=skeys,
                       z32=z32[:m] if z32 is not None else None, chunk_minmax=mm)


def _voxel_downsample_wide(dl: DeviceLas, voxel_size: float, cs: int, want, sink) -> VoxelResult:
    """The voxel index range does not fit one sort word: decode to float64 and take the wide-key path."""
    pts = decode_xyz(dl, torch.float64)
    w = voxel_downsample_points(pts, voxel_size, cs)
    lat = quantise(w.mean, dl.scales, dl.offsets) if (sink is not None or {"lattice", "f32", "z32"} & set(want)) else None
    f32 = None
    if "f32" in want or "z32" in want or sink is not None:
        # astype(float32) of the re-quantised values: encode the lattice into minimal records and run the
        # ordinary float32 decode kernel over them
        recs, _ = encode_records(lat, 20)
        f32 = decode_xyz(DeviceLas(recs, w.count, 20, dl.scales, dl.offsets), torch.float32)
    if sink is not None:
        if sink.f32.shape[0] - sink.count < w.count:
            raise ValueError("VoxelSink too small for this slice")
        sink.f32[sink.count: sink.count + w.count].copy_(f32)
        if sink.z32 is not None:
            sink.z32[sink.count: sink.count + w.count].copy_(f32[:, 2])
        sink.count += w.count
    return VoxelResult(w.count, w.chunk_counts, w.mean if "mean" in want else None,
                       lat if "lattice" in want else None, f32, plan=w.plan,
                       z32=f32[:, 2].contiguous() if "z32" in want else None)


# ------------------------------------------------------------------------------------------------
# tower extraction, device stages
# ------------------------------------------------------------------------------------------------
def f32_centroid(xyz: torch.Tensor, serial: bool = False, want_stats: bool = False):
    """np.mean(raw_points_f32, axis=0) bit-exactly: (centroid float32[3], sequential sums float32[3]).
    serial=True runs the one-thread-per-column reference kernel instead of the parallel evaluation."""
    assert xyz.dtype == torch.float32 and xyz.is_contiguous()
    lib = _native.lib()
    sums = torch.empty(3, dtype=torch.float32, device=xyz.device)
    cen = torch.empty(3, dtype=torch.float32, device=xyz.device)
    ws, wsb = None, 0
    if not serial:
        wsb = lib.pch_f32_centroid_workspace_bytes(xyz.shape[0])
        ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
    check(lib.pch_f32_centroid(xyz.data_ptr(), xyz.shape[0], sums.data_ptr(), cen.data_ptr(), _ptr(ws), wsb,
                               _stream()), "pch_f32_centroid")
    if want_stats:
        st = ws[:24].view(torch.int32).cpu().numpy().reshape(3, 2) if ws is not None else None
        return cen, sums, st
    return cen, sums


def f32_shift(xyz: torch.Tensor, centroid: torch.Tensor, want_z=True, want_xyz=False):
    m = xyz.shape[0]
    zs = torch.empty(m, dtype=torch.float32, device=xyz.device) if want_z else None
    sh = torch.empty((m, 3), dtype=torch.float32, device=xyz.device) if want_xyz else None
    check(_native.lib().pch_f32_shift(xyz.data_ptr(), m, centroid.data_ptr(), _ptr(zs), _ptr(sh), _stream()),
          "pch_f32_shift")
    return zs, sh


def f32_column(xyz: torch.Tensor, column: int) -> torch.Tensor:
    """xyz[:, column] as a dense (m,) float32 array."""
    assert xyz.dtype == torch.float32 and xyz.is_contiguous()
    out = torch.empty(xyz.shape[0], dtype=torch.float32, device=xyz.device)
    check(_native.lib().pch_f32_column(xyz.data_ptr(), xyz.shape[0], int(column), out.data_ptr(), _stream()),
          "pch_f32_column")
    return out


def select_f32(v: torch.Tensor, rank0: int, rank1: int) -> torch.Tensor:
    """Exact order statistics sorted(v)[rank0], sorted(v)[rank1] as a float32[2] device tensor."""
    assert v.dtype == torch.float32 and v.is_contiguous()
    lib = _native.lib()
    wsb = lib.pch_select_workspace_bytes()
    ws = torch.empty(wsb, dtype=torch.uint8, device=v.device)
    out = torch.empty(2, dtype=torch.float32, device=v.device)
    check(lib.pch_select_f32(v.data_ptr(), v.numel(), int(rank0), int(rank1), out.data_ptr(), ws.data_ptr(), wsb,
                             _stream()), "pch_select_f32")
    return out


def compact_points(xyz: torch.Tensor, zs: Optional[torch.Tensor], thr: float,
                   centroid: Optional[torch.Tensor] = None, keep_mask: Optional[torch.Tensor] = None,
                   want_src: bool = False, want_mask: bool = False):
    """(filtered (G,3) float32 = xyz[keep] - centroid, G, src index or None, mask or None).
    keep = zs > thr, or keep_mask != 0, or (both None) float32(xyz.z - centroid.z) > thr on the fly."""
    if zs is None and keep_mask is None and centroid is None:
        raise ValueError("compact_points needs zs, keep_mask or a centroid")
    assert xyz.dtype == torch.float32 and xyz.is_contiguous()
    lib = _native.lib()
    m = xyz.shape[0]
    dev = xyz.device
    out = torch.empty((m, 3), dtype=torch.float32, device=dev)
    src = torch.empty(m, dtype=torch.int32, device=dev) if want_src else None
    mask = torch.empty(m, dtype=torch.uint8, device=dev) if want_mask else None
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    wsb = lib.pch_compact_workspace_bytes(m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.pch_compact_points(xyz.data_ptr(), _ptr(zs), _ptr(keep_mask), m, _ptr(centroid), float(thr),
                                 out.data_ptr(), _ptr(src), _ptr(mask), cnt.data_ptr(), ws.data_ptr(), wsb,
                                 _stream()), "pch_compact_points")
    g = int(cnt.item())
    return out[:g], g, (src[:g] if src is not None else None), mask


def f32_minmax(xyz: torch.Tensor) -> np.ndarray:
    out = torch.empty(6, dtype=torch.float32, device=xyz.device)
    check(_native.lib().pch_f32_minmax(xyz.data_ptr(), xyz.shape[0], out.data_ptr(), _stream()), "pch_f32_minmax")
    return out.cpu().numpy()


GRID_MAX_CELLS_PER_POINT = 64     # a cell table beyond max(this * m, 2^22) cells means an outlier blew the extent up


def grid_shape(mn_xy, mx_xy, cell: float, m: int):
    """(nx, ny) of the grid over [mn, mx] (float32 arithmetic, like the kernels); ValueError when the table would
    be out of proportion to the cloud (one far outlier) or overflow the int32 cell index of the C ABI."""
    c = np.float32(cell)
    if not (c > 0):
        raise ValueError("cell must be > 0")
    mn = np.asarray(mn_xy, dtype=np.float32)
    mx = np.asarray(mx_xy, dtype=np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        ext = np.floor((mx - mn) / c).astype(np.float64)
    if not np.all(np.isfinite(ext)):
        raise ValueError("grid extent is not finite")
    nx, ny = int(ext[0]) + 1, int(ext[1]) + 1
    cap = max(GRID_MAX_CELLS_PER_POINT * max(int(m), 1), 1 << 22)
    if nx * ny > min(cap, 2 ** 31 - 1):
        raise ValueError(f"grid of {nx} x {ny} cells for {m} points: extent out of proportion (outlier?) — crop the cloud "
                         f"or use a larger cell")
    return nx, ny


def grid_min_ground(xyz: torch.Tensor, cell: float = 2.0, hag: float = 3.0):
    """north_star grid min-z ground model on an already shifted cloud: (keep mask uint8 (m), ground_z float32 (m))."""
    assert xyz.dtype == torch.float32 and xyz.is_contiguous()
    m = xyz.shape[0]
    dev = xyz.device
    if m == 0:
        return torch.zeros(0, dtype=torch.uint8, device=dev), torch.zeros(0, dtype=torch.float32, device=dev)
    mm = f32_minmax(xyz)
    c = np.float32(cell)
    nx, ny = grid_shape(mm[0:2], mm[3:5], cell, m)
    cell_min = torch.empty(nx * ny, dtype=torch.int32, device=dev)
    keep = torch.empty(m, dtype=torch.uint8, device=dev)
    gz = torch.empty(m, dtype=torch.float32, device=dev)
    check(_native.lib().pch_grid_min_ground(xyz.data_ptr(), m, float(mm[0]), float(mm[1]), float(c), nx, ny,
                                            float(np.float32(hag)), cell_min.data_ptr(), keep.data_ptr(),
                                            gz.data_ptr(), _stream()), "pch_grid_min_ground")
    return keep, gz


def f32_minmax_dev(xyz: torch.Tensor) -> torch.Tensor:
    out = torch.empty(6, dtype=torch.float32, device=xyz.device)
    check(_native.lib().pch_f32_minmax(xyz.data_ptr(), xyz.shape[0], out.data_ptr(), _stream()), "pch_f32_minmax")
    return out


def compact_points_grid(raw: torch.Tensor, centroid: torch.Tensor, mn_xy, nx: int, ny: int, cell: float, hag: float,
                        want_mask: bool = False, sync: bool = True):
    """Grid min-z filter of the RAW cloud with the centroid shift applied on the fly: (filtered (G,3) float32 =
    (raw - centroid)[keep], G, mask or None).  Two kernels: the cell minima, then the compaction whose keep flag is
    the height-above-ground test."""
    assert raw.dtype == torch.float32 and raw.is_contiguous()
    lib = _native.lib()
    m = raw.shape[0]
    dev = raw.device
    st = _stream()
    c = float(np.float32(cell))
    cell_min = torch.empty(nx * ny, dtype=torch.int32, device=dev)
    check(lib.pch_grid_min(raw.data_ptr(), m, centroid.data_ptr(), float(mn_xy[0]), float(mn_xy[1]), c, nx, ny,
                           cell_min.data_ptr(), st), "pch_grid_min")
    out = torch.empty((m, 3), dtype=torch.float32, device=dev)
    mask = torch.empty(m, dtype=torch.uint8, device=dev) if want_mask else None
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    wsb = lib.pch_compact_workspace_bytes(m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.pch_compact_points_grid(raw.data_ptr(), m, centroid.data_ptr(), float(mn_xy[0]), float(mn_xy[1]), c, nx, ny,
                                      float(np.float32(hag)), cell_min.data_ptr(), out.data_ptr(), None, _ptr(mask),
                                      cnt.data_ptr(), ws.data_ptr(), wsb, st), "pch_compact_points_grid")
    if not sync:
        return out, cnt, mask            # the caller reads the count (with others) and slices out[:count]
    g = int(cnt.item())
    return out[:g], g, mask


@dataclasses.dataclass
class DbscanResult:
    labels: torch.Tensor          # int32 [G], the reference's all_labels
    n_clusters: int
    stats: np.ndarray             # structured host array [K]: count, min[3], max[3], sum[3]
    plan: Optional[dict] = None   # cell-grid key layout (bits per axis, radix passes)


DB_STATS_PEEK = 4096     # rows of the per-cluster table fetched together with the scalar block
STATS_DTYPE = np.dtype([("count", "<i8"), ("min", "<f4", 3), ("max", "<f4", 3), ("sum", "<f8", 3)])
assert STATS_DTYPE.itemsize == C.sizeof(_native.ClusterStats) == 56


def dbscan_chunked(points: torch.Tensor, eps: float = 8.0, min_samples: int = 80, chunk: int = 50000) -> DbscanResult:
    """Chunked sklearn-exact DBSCAN of the (G,3) float32 candidates (utils/tower_extraction.py:96-122)."""
    _require_cuda()
    assert points.dtype == torch.float32 and points.is_contiguous()
    lib = _native.lib()
    dev = points.device
    G = points.shape[0]
    if G == 0:
        return DbscanResult(torch.zeros(0, dtype=torch.int32, device=dev), 0, np.zeros(0, dtype=STATS_DTYPE))
    ch = max(1, min(int(chunk), G))
    n_chunks = -(-G // ch)
    st = _stream()
    labels = torch.empty(G, dtype=torch.int32, device=dev)
    cap = max(DB_STATS_PEEK, G // 256)
    item = STATS_DTYPE.itemsize
    while True:
        stats = torch.empty(cap * item, dtype=torch.uint8, device=dev)
        wsb = lib.pch_dbscan_fused_workspace_bytes(G, ch, cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        # one call, no host round trip inside (the cell-grid plan stays on the device)
        check(lib.pch_dbscan(points.data_ptr(), G, ch, float(eps), int(min_samples), labels.data_ptr(), stats.data_ptr(), cap,
                             ws.data_ptr(), wsb, st), "pch_dbscan")
        # ONE device->host read: the scalar block and the first DB_STATS_PEEK rows of the per-cluster table
        peek = min(cap, DB_STATS_PEEK)
        both = torch.cat([ws[:256], stats[: peek * item]]).cpu().numpy()
        sc = both[:256]
        k = int(sc[136:144].view(np.int64)[0])
        plan = VoxelPlan(*[int(v) for v in sc[208:240].view(np.int32)])
        if plan.status != 0:
            raise ValueError("DBSCAN cell grid does not fit the packed key")
        if os.environ.get("PCH_TRACE"):
            print(f"[pch] dbscan G={G} chunks={n_chunks} cells={int(sc[128:136].view(np.int64)[0])} "
                  f"non-dense points={int(sc[192:196].view(np.uint32)[0])} clusters={k} plan={plan.bits_x},{plan.bits_y},{plan.bits_z}",
                  flush=True)
        if int(sc[:4].view(np.int32)[0]):
            raise _native.NativeError("device look-back spin limit hit in dbscan")
        if k <= cap:
            break
        cap = k
    if k <= peek:
        host = both[256: 256 + k * item].view(STATS_DTYPE).copy()
    else:
        host = stats[: k * item].cpu().numpy().view(STATS_DTYPE).copy()
    return DbscanResult(labels, k, host, {f: getattr(plan, f) for f, _ in VoxelPlan._fields_})


def voxel_downsample_points(xyz: torch.Tensor, voxel_size: float, chunk_size: Optional[int] = None) -> VoxelResult:
    """open3d voxel_down_sample of an arbitrary (n,3) float64 device array (process_chunk's input)."""
    _require_cuda()
    assert xyz.dtype == torch.float64 and xyz.is_contiguous() and xyz.dim() == 2 and xyz.shape[1] == 3
    if voxel_size <= 0:
        raise ValueError("voxel_size must be > 0")
    lib = _native.lib()
    dev = xyz.device
    n = xyz.shape[0]
    cs = max(1, min(int(chunk_size or n or 1), max(n, 1)))
    n_chunks = max(1, -(-n // cs))
    if n == 0:
        return VoxelResult(0, torch.zeros(n_chunks, dtype=torch.int64, device=dev),
                           torch.zeros((0, 3), dtype=torch.float64, device=dev))
    st = _stream()
    scratch = torch.empty(n_chunks * 6, dtype=torch.int64, device=dev)
    origins = torch.empty((n_chunks, 3), dtype=torch.float64, device=dev)
    plan_dev = torch.empty(8, dtype=torch.int32, device=dev)
    check(lib.pch_voxel_plan_build_f64(xyz.data_ptr(), n, cs, float(voxel_size), scratch.data_ptr(), origins.data_ptr(),
                                       plan_dev.data_ptr(), st), "pch_voxel_plan_build_f64")
    plan = VoxelPlan(*[int(v) for v in plan_dev.cpu().numpy()])
    if max(plan.bits_x, plan.bits_y, plan.bits_z) > 31:
        raise ValueError("voxel_size is too small.")          # open3d: index does not fit an int
    vidx = None
    if plan.status != 0:
        # wide keys: three stable sort rounds (z, y, x) over  axis_index << bits_idx | index  words
        vidx = torch.empty((n, 3), dtype=torch.int32, device=dev)
        check(lib.pch_voxel_index3_f64(xyz.data_ptr(), n, cs, float(voxel_size), origins.data_ptr(), vidx.data_ptr(), st),
              "pch_voxel_index3_f64")
        skeys = None
        for axis, bits in ((2, plan.bits_z), (1, plan.bits_y), (0, plan.bits_x)):
            words = torch.empty(n, dtype=torch.int64, device=dev)
            check(lib.pch_voxel_wide_words(_ptr(skeys), vidx.data_ptr(), n, cs, plan.bits_idx, axis, words.data_ptr(), st),
                  "pch_voxel_wide_words")
            skeys = sort_u64_segmented(words, cs, plan.bits_idx, plan.bits_idx + max(bits, 1))
    else:
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        check(lib.pch_voxel_keys_f64(xyz.data_ptr(), n, cs, float(voxel_size), origins.data_ptr(), C.byref(plan),
                                     keys.data_ptr(), st), "pch_voxel_keys_f64")
        skeys = sort_u64_segmented(keys, cs, plan.bits_idx, plan.bits_idx + plan.key_bits)
    mean = torch.empty((n, 3), dtype=torch.float64, device=dev)
    counts = torch.empty(n_chunks, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    wsb = lib.pch_voxel_reduce_workspace_bytes(n, cs)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.pch_voxel_reduce(skeys.data_ptr(), n, cs, plan.bits_idx, xyz.data_ptr(), 0, None, _ptr(vidx), None, None,
                               mean.data_ptr(), None, None, None, counts.data_ptr(), total.data_ptr(), ws.data_ptr(), wsb,
                               st), "pch_voxel_reduce")
    m = int(total.item())
    return VoxelResult(m, counts, mean[:m], plan={k: getattr(plan, k) for k, _ in VoxelPlan._fields_})


class ObbResult(C.Structure):
    _fields_ = [("extents", C.c_double * 3), ("center", C.c_double * 3), ("rotation", C.c_double * 9), ("volume", C.c_double),
                ("n_faces", C.c_int32), ("n_vertices", C.c_int32), ("n_candidates", C.c_int32), ("status", C.c_int32)]


OBB_DTYPE = np.dtype([("extents", "<f8", 3), ("center", "<f8", 3), ("rotation", "<f8", (3, 3)), ("volume", "<f8"),
                      ("n_faces", "<i4"), ("n_vertices", "<i4"), ("n_candidates", "<i4"), ("status", "<i4")])
assert OBB_DTYPE.itemsize == C.sizeof(ObbResult) == 144


def obb_batch(rows: torch.Tensor, ranges: np.ndarray) -> np.ndarray:
    """Minimum-volume oriented boxes of a batch of clusters on the device (pch_obb_batch): `rows` (L,3) float32,
    `ranges` int64 (K,2) = first / end row of each cluster.  -> structured host array [K] (OBB_DTYPE)."""
    _require_cuda()
    assert rows.dtype == torch.float32 and rows.is_contiguous()
    rg = np.ascontiguousarray(ranges, dtype=np.int64).reshape(-1, 2)
    K = rg.shape[0]
    if K == 0:
        return np.zeros(0, dtype=OBB_DTYPE)
    lib = _native.lib()
    dev = rows.device
    rg_dev = torch.from_numpy(rg).to(dev)
    out = torch.empty(K * OBB_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    wsb = lib.pch_obb_workspace_bytes(K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.pch_obb_batch(rows.data_ptr(), rg_dev.data_ptr(), K, out.data_ptr(), ws.data_ptr(), wsb, _stream()), "pch_obb_batch")
    return out.cpu().numpy().view(OBB_DTYPE).copy()


def cluster_major_points(points: torch.Tensor, labels: torch.Tensor, counts: np.ndarray):
    """All `points[labels == k]` at once: ((L,3) float32 rows grouped by label in ascending label order, each
    group in original order; int64 offsets [K+1]).  L = number of labelled points."""
    assert points.dtype == torch.float32 and points.is_contiguous() and labels.dtype == torch.int32
    lib = _native.lib()
    G = labels.numel()
    K = len(counts)
    offsets = np.zeros(K + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    L = int(offsets[-1])
    out = torch.empty((L, 3), dtype=torch.float32, device=points.device)
    if L == 0:
        return out, offsets
    words = torch.empty(G, dtype=torch.int64, device=points.device)
    check(lib.pch_label_words(labels.data_ptr(), G, words.data_ptr(), _stream()), "pch_label_words")
    bits = max(1, int(K).bit_length())
    sw = sort_u64_segmented(words, G, 32, 32 + min(31, bits))
    check(lib.pch_gather_rows_f32(points.data_ptr(), sw.data_ptr(), L, out.data_ptr(), None, _stream()),
          "pch_gather_rows_f32")
    return out, offsets
